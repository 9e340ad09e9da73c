"""Golden vectors for the training losses, produced by the UNMODIFIED reference
(trajnetbaselines/lstm/loss.py: PredictionLoss :6-91, L2Loss :93-135, CollisionLoss :138-162)
imported from /root/reference in the build container.

    python -m oracle.make_loss_golden        -> tests/golden/loss_golden.npz

Cases: ragged scenes (incl. a single-pedestrian scene), NaN neighbours, neighbours placed inside
and exactly outside the collision radius; value and the gradients wrt the network outputs
(inputs [T, M, 5]) and wrt the positions handed to the collision term.
"""
import os

import numpy as np


def make_case(seed, sizes, T=12):
    rng = np.random.RandomState(seed)
    bs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    M = int(bs[-1])
    inputs = np.zeros((T, M, 5), np.float32)
    inputs[..., :2] = rng.randn(T, M, 2) * 0.3
    inputs[..., 2:4] = 0.01 + 0.2 / (1 + np.exp(-rng.randn(T, M, 2)))      # Hidden2Normal ranges
    inputs[..., 4] = 0.7 / (1 + np.exp(-rng.randn(T, M))) * np.sign(rng.randn(T, M))
    targets = (rng.randn(T, M, 2) * 0.3).astype(np.float32)
    pos = (rng.randn(T, M, 2) * 1.5).astype(np.float32)
    for b in range(len(sizes)):                 # put some neighbours close to their primary
        a, e = bs[b], bs[b + 1]
        for j in range(a + 1, e):
            if rng.rand() < 0.5:
                t = rng.randint(T)
                ang = rng.rand() * 6.28
                r = rng.choice([0.05, 0.12, 0.19, 0.21])
                pos[t, j] = pos[t, a] + r * np.array([np.cos(ang), np.sin(ang)], np.float32)
    nan_rows = [int(bs[b] + 1) for b in range(len(sizes)) if sizes[b] > 2][:2]
    for j in nan_rows:
        pos[T // 2:, j] = np.nan
    return inputs, targets, pos, bs


def reference_outputs(inputs, targets, pos, bs, col_wt):
    import torch
    from oracle.ref_shim import import_reference
    import_reference()
    from trajnetbaselines.lstm.loss import L2Loss, PredictionLoss
    out = {}
    for name, cls in (("pl", PredictionLoss), ("l2", L2Loss)):
        i = torch.from_numpy(inputs.copy()).requires_grad_(True)
        p = torch.from_numpy(pos.copy()).requires_grad_(True)
        crit = cls(col_wt=col_wt, col_distance=0.2)
        loss = crit(i, torch.from_numpy(targets), torch.from_numpy(bs), (p * 1.0) if col_wt else None)
        loss.backward()
        out[name + "/loss"] = np.array([loss.item()], np.float64)
        out[name + "/dinputs"] = i.grad.numpy().copy()
        out[name + "/dpos"] = p.grad.numpy().copy() if p.grad is not None else np.zeros_like(pos)
    i = torch.from_numpy(inputs.copy())
    out["pl/keep_batch"] = PredictionLoss(keep_batch_dim=True)(i, torch.from_numpy(targets), torch.from_numpy(bs)).numpy()
    return out


CASES = {"uniform": (1, [5, 5, 5, 5]), "ragged": (2, [1, 7, 2, 12, 3]), "big": (3, [20] * 6)}


def main():
    blob = {}
    for name, (seed, sizes) in CASES.items():
        inputs, targets, pos, bs = make_case(seed, sizes)
        blob[name + "/inputs"], blob[name + "/targets"], blob[name + "/pos"], blob[name + "/bs"] = inputs, targets, pos, bs
        for col_wt in (0.0, 10.0):
            for k, v in reference_outputs(inputs, targets, pos, bs, col_wt).items():
                blob["%s/col%d/%s" % (name, int(col_wt), k)] = v
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "loss_golden.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, len(blob), "arrays")


if __name__ == "__main__":
    main()
