"""Search for the smallest scene on which the CUDA social backward disagrees with the reference autograd, and
print the pair structure (cells per observer) of every step for it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lstm_oracle as O
from oracle.ref_shim import import_reference
import_reference()
from oracle.make_golden import build_reference_model
from trajnetbaselines.lstm.loss import PredictionLoss as RefLoss
from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling, PredictionLoss
torch.set_num_threads(4)
kind = "social_small"
W = O.random_weights(kind, seed=11)
ref = build_reference_model(kind, W); ref.train()
mine = LSTM(pool=GridBasedPooling(**O.MODEL_SPECS[kind])); mine.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()}); mine = mine.cuda().train()

def run(xy, bs):
    scene = torch.from_numpy(xy); split = torch.from_numpy(bs); M = xy.shape[1]
    out = {}
    for name, model, crit, dev in (("ref", ref, RefLoss(), "cpu"), ("b200", mine, PredictionLoss(), "cuda")):
        s = scene.to(dev)
        rel, pred = model(s[:9], torch.zeros(M, 2).to(dev), split.to(dev), s[9:-1].clone())
        loss = crit(rel[-12:], s[9:21] - s[8:20], split.to(dev)) * (len(bs) - 1)
        model.zero_grad(); loss.backward()
        out[name] = (model.pool.hidden_dim_encoding.bias.grad.detach().cpu().numpy().copy(), pred.detach().cpu().numpy())
    g, gb = out["ref"][0], out["b200"][0]
    return float(np.abs(gb - g).max() / max(np.abs(g).max(), 1e-12)), g, gb, out["ref"][1]

def run_steps(xy, bs):
    """per-step d lat_j: reference via a hook on hidden_dim_encoding's output, CUDA via TB2_DUMP_DLAT."""
    scene = torch.from_numpy(xy); split = torch.from_numpy(bs); M = xy.shape[1]; N = M
    grads_ref = []
    def fwd_hook(mod, inp, out):
        new = out * 1.0
        new.register_hook(lambda g: grads_ref.append(g.detach().clone()))
        return new
    h = ref.pool.hidden_dim_encoding.register_forward_hook(fwd_hook)
    rel, pred = ref(scene[:9], torch.zeros(M, 2), split, scene[9:-1].clone())
    loss = RefLoss()(rel[-12:], scene[9:21] - scene[8:20], split) * 1
    ref.zero_grad(); loss.backward(); h.remove()
    grads_ref = grads_ref[::-1]            # hooks fire in reverse step order
    # [1, N, N-1, C] per step -> d lat_j = sum over observers i of the entry (i, jj(j))
    dl_ref = np.zeros((len(grads_ref), N, grads_ref[0].shape[-1]), np.float32)
    for s, g in enumerate(grads_ref):
        g = g.numpy()[0]
        for i in range(N):
            for jj in range(N - 1):
                j = jj + (jj >= i)
                dl_ref[s, j] += g[i, jj]
    os.environ["TB2_DUMP_DLAT"] = "/tmp/dlat.bin"
    s_ = scene.cuda()
    rel, _ = mine(s_[:9], torch.zeros(M, 2).cuda(), split.cuda(), s_[9:-1].clone())
    loss = PredictionLoss()(rel[-12:], s_[9:21] - s_[8:20], split.cuda()) * 1
    mine.zero_grad(); loss.backward(); torch.cuda.synchronize()
    dl = np.fromfile("/tmp/dlat.bin", dtype=np.float32).reshape(-1, M, dl_ref.shape[-1])
    del os.environ["TB2_DUMP_DLAT"]
    return dl_ref, dl, pred.detach().numpy()

bad = []
for N in (4,):
    for seed in range(3):
        xy, bs = O.synthetic_scenes(1, N, seed=100 + seed, start_std=1.2)
        err, g, gb, pred = run(xy, bs)
        print("N=%d seed=%d  rel err db_enc %.3e" % (N, seed, err), flush=True)
        if err > 1e-3:
            bad.append((N, seed, err))
print("failing:", bad)
if bad:
    N, seed, err = bad[0]
    xy, bs = O.synthetic_scenes(1, N, seed=100 + seed, start_std=1.2)
    err, g, gb, pred = run(xy, bs)
    print("smallest failing case N=%d seed=%d err %.3e" % (N, seed, err))
    print("db_enc ref ", g)
    print("db_enc b200", gb)
    dl_ref, dl, _ = run_steps(xy, bs)
    print("steps ref %d cuda %d" % (len(dl_ref), len(dl)))
    for s in range(min(len(dl_ref), len(dl))):
        d = np.abs(dl[s] - dl_ref[s]).max(axis=1)
        sc = max(np.abs(dl_ref[s]).max(), 1e-12)
        print("  step %2d  max|dlat - ref| per track: %s   (scale %.2e)   ref row0 %s cuda row0 %s" % (
            s, np.array2string(d / sc, precision=3), sc, np.array2string(dl_ref[s][:, 0], precision=4), np.array2string(dl[s][:, 0], precision=4)))
    cfg = O.pool_config(kind)
    # step inputs of the teacher-forced pass: obs2 of step s (primary rows from the predictions for the decoder)
    S = 19
    for s in range(S):
        if s < 8:
            obs2 = xy[s + 1].copy()
        else:
            obs2 = xy[s + 1].copy() if s + 1 < 21 else None
            obs2[0] = pred[s - 1][0]          # primary overwritten by the (detached) prediction of the previous step
        rel = obs2[None, :, :] - obs2[:, None, :]
        oij = rel / cfg.cell_side + cfg.n / 2
        rows = []
        for i in range(N):
            ent = []
            for j in range(N):
                if j == i: continue
                o = oij[i, j]
                inr = (o >= 0).all() and (o < cfg.n).all()
                ent.append("%d:%s" % (j, ("c%d" % (int(o[0]) * cfg.n + int(o[1]))) if inr else "out"))
            rows.append("i%d[%s]" % (i, " ".join(ent)))
        print("step %2d  %s" % (s, "  ".join(rows)))
