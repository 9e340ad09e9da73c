"""Golden vectors of the non-grid interaction modules HiddenStateMLPPooling, NearestNeighborMLP and AttentionMLPPooling, produced
by the UNMODIFIED reference (trajnetbaselines/lstm/non_gridbased_pooling.py:64-147,150-239,242-351 inside trajnetbaselines.lstm.LSTM).

TEST INFRASTRUCTURE.  Run in the build container (needs the reference): python -m oracle.make_nongrid_golden
-> tests/golden/nongrid_golden.npz.  Cases: the stand-alone plug on padded scenes with NaN tracks, and
LSTM.forward free-running and teacher-forced on ragged scenes with late / early tracks."""
import os

import numpy as np
import torch

from oracle import lstm_oracle as O
from oracle.ref_shim import import_reference

KINDS = ["hiddenstatemlp", "hiddenstatemlp_small"]
NN_KINDS = ["nn", "nn_small"]
ATTN_KINDS = ["attentionmlp", "attentionmlp_small"]
NN_LSTM_KINDS = ["nn_lstm", "nn_lstm_small"]
TRAJ_KINDS = ["traj_pool", "traj_pool_small"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plug_inputs():
    rng = np.random.RandomState(5)
    B, N = 6, 9
    obs2 = (rng.randn(B, N, 2) * 2.0).astype(np.float32)
    obs1 = obs2 - (rng.randn(B, N, 2) * 0.3).astype(np.float32)
    hid = (rng.randn(B, N, 128) * 0.5).astype(np.float32)
    obs2[1, 4:] = np.nan          # padded scene
    obs1[1, 4:] = np.nan
    hid[1, 4:] = np.nan
    obs1[2, 3] = np.nan           # present now, absent before (relative velocity NaN)
    obs2[3, 5] = np.nan           # absent now
    return hid, obs1, obs2


def scene_inputs():
    return O.synthetic_scenes(9, 8, seed=29, ragged=True, nan_tracks=True)


def build_reference_model(kind, W):
    from trajnetbaselines.lstm import LSTM
    from trajnetbaselines.lstm.non_gridbased_pooling import (AttentionMLPPooling, HiddenStateMLPPooling, NearestNeighborLSTM,
                                                              NearestNeighborMLP, TrajectronPooling)
    if kind in O.TRAJ_SPECS:
        pool = TrajectronPooling(**O.TRAJ_SPECS[kind])
    elif kind in O.NN_LSTM_SPECS:
        pool = NearestNeighborLSTM(**O.NN_LSTM_SPECS[kind])
    elif kind in O.NN_SPECS:
        pool = NearestNeighborMLP(**O.NN_SPECS[kind])
    elif kind in O.ATTN_SPECS:
        pool = AttentionMLPPooling(**O.ATTN_SPECS[kind])
    else:
        pool = HiddenStateMLPPooling(**O.NONGRID_SPECS[kind])
    model = LSTM(pool=pool)
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()}, strict=True)
    return model.eval()


def main():
    import_reference()
    out = {}
    hid, obs1, obs2 = plug_inputs()
    xy, bs = scene_inputs()
    M = xy.shape[1]
    for kind in KINDS + NN_KINDS + ATTN_KINDS + NN_LSTM_KINDS + TRAJ_KINDS:
        W = O.random_weights(kind, seed=13)
        model = build_reference_model(kind, W)
        with torch.no_grad():
            if kind in NN_LSTM_KINDS + TRAJ_KINDS:     # stateful plug: two consecutive calls after a reset
                model.pool.reset(obs2.shape[0] * obs2.shape[1], obs2.shape[1] - 1, device=torch.device("cpu"))
                first = model.pool(torch.from_numpy(hid), torch.from_numpy(obs1), torch.from_numpy(obs2)).numpy()
                second = model.pool(torch.from_numpy(hid), torch.from_numpy(obs2), torch.from_numpy(obs2 + (obs2 - obs1))).numpy()
                out[kind + "/plug"] = first
                out[kind + "/plug2"] = second
            else:
                out[kind + "/plug"] = model.pool(torch.from_numpy(hid), torch.from_numpy(obs1), torch.from_numpy(obs2)).numpy()
            rel, pred = model(torch.from_numpy(xy[:9]), torch.zeros(M, 2), torch.from_numpy(bs), n_predict=12)
            rel_t, pred_t = model(torch.from_numpy(xy[:9]), torch.zeros(M, 2), torch.from_numpy(bs),
                                  prediction_truth=torch.from_numpy(xy[9:20]).clone())
        out[kind + "/rel_free"] = rel.numpy()
        out[kind + "/pred_free"] = pred.numpy()
        out[kind + "/pred_teacher"] = pred_t.numpy()
    path = os.path.join(ROOT, "tests", "golden", "nongrid_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
