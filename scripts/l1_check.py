"""A/B check of the first grid-embedding Linear of the social pool: TB2_SPARSE = tc (round-1 tcgen05
kernel) / solo (round-2 kernel, one CTA) / pair (round-2 kernel, CTA pair) against the numpy oracle (small
case) and against each other (BASELINE-size case), with CUDA-event timing of the pool call.

    python scripts/l1_check.py --mode pair
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="pair")
    ap.add_argument("--scenes", type=int, default=256)
    ap.add_argument("--iters", type=int, default=70)
    args = ap.parse_args()
    os.environ["TB2_SPARSE"] = args.mode
    import torch
    from oracle import lstm_oracle as O
    from trajnetplusplusbaselines_b200.lstm import GridBasedPooling

    O.MODEL_SPECS["social_l1"] = dict(type_="social", hidden_dim=128, cell_side=0.6, n=16, out_dim=1024,
                                      embedding_arch="one_layer", latent_dim=16)

    def make_pool(kind, W):
        pool = GridBasedPooling(**O.MODEL_SPECS[kind])
        sd = {k[len("pool."):]: torch.from_numpy(v.copy()) for k, v in W.items() if k.startswith("pool.")}
        pool.load_state_dict(sd, strict=True)
        return pool.cuda()

    # ---- small case vs oracle (layer-1 output directly, then the full two_layer pool) ----------
    rng = np.random.RandomState(5)
    B, N = 9, 11
    obs2 = (rng.randn(B, N, 2) * 2.0).astype(np.float32)
    obs1 = obs2 - (rng.randn(B, N, 2) * 0.3).astype(np.float32)
    hid = (rng.randn(B, N, 128) * 0.5).astype(np.float32)
    obs2[1, 4:] = np.nan
    obs1[1, 4:] = np.nan
    hid[1, 4:] = np.nan
    obs1[2, 3] = np.nan
    obs2[3, 5] = np.nan
    for kind in ("social_l1", "social"):
        cfg = O.pool_config(kind)
        W = O.random_weights(kind, seed=21)
        ref = O.pool_forward(cfg, W, hid, obs1, obs2)
        pool = make_pool(kind, W)
        out = pool(torch.from_numpy(hid).cuda(), torch.from_numpy(obs1).cuda(), torch.from_numpy(obs2).cuda())
        torch.cuda.synchronize()
        out = out.cpu().numpy()
        d = np.abs(out - ref)
        print("[%s] small %-10s max|gpu - oracle| = %.3e  (ref max %.3f, nan mismatch %d)" % (
            args.mode, kind, float(np.nanmax(d)), float(np.nanmax(np.abs(ref))),
            int((np.isnan(out) != np.isnan(ref)).sum())), flush=True)
        if np.nanmax(d) > 1e-4:
            bad = np.argwhere(d > 1e-4)
            print("   first bad entries (row, col):", bad[:8].tolist(), " count", len(bad), "of", d.size)
            rows = np.unique(bad[:, 0])
            print("   bad rows:", rows[:40].tolist())
            cols = np.unique(bad[:, 1])
            print("   bad cols (first 40):", cols[:40].tolist(), " n bad cols", len(cols))

    # ---- BASELINE-size case: layer-1 output of the mode vs the SS-form kernel, ragged too -------
    for ragged in (False, True):
        xy, bs = O.synthetic_scenes(args.scenes, 20, n_frames=4, seed=7, ragged=ragged, nan_tracks=ragged)
        M = xy.shape[1]
        rng = np.random.RandomState(11)
        h = (rng.randn(M, 128) * 0.5).astype(np.float32)
        sizes = np.diff(bs)
        nmax = int(sizes.max())
        # the pool plug takes padded [B, Nmax, .] tensors
        def pad(a, fill):
            out = np.full((len(sizes), nmax) + a.shape[1:], fill, dtype=np.float32)
            for b in range(len(sizes)):
                out[b, :sizes[b]] = a[bs[b]:bs[b + 1]]
            return out
        o1, o2, hh = pad(xy[2], np.nan), pad(xy[3], np.nan), pad(h, np.nan)
        W = O.random_weights("social_l1", seed=21)
        pool = make_pool("social_l1", W)
        t_in = [torch.from_numpy(a).cuda() for a in (hh, o1, o2)]
        os.environ["TB2_SPARSE"] = "tc"
        ref = pool(*t_in).cpu().numpy()
        os.environ["TB2_SPARSE"] = args.mode
        out = pool(*t_in)
        torch.cuda.synchronize()
        out = out.cpu().numpy()
        d = np.abs(out - ref)
        print("[%s] big ragged=%d M=%d  max|mode - tc| = %.3e  mean %.3e  (ref max %.3f)" % (
            args.mode, ragged, M, float(np.nanmax(d)), float(np.nanmean(d)), float(np.nanmax(np.abs(ref)))), flush=True)
        if np.nanmax(d) > 1e-4:
            bad = np.argwhere(d > 1e-4)
            print("   bad count", len(bad), "of", d.size, " rows", np.unique(bad[:, 0])[:20].tolist(),
                  " n bad rows", len(np.unique(bad[:, 0])), " n bad cols", len(np.unique(bad[:, 1])))
        # the full two_layer pool (with the round-2 kernel the 1024 -> 256 Linear runs inside it): vs the round-1 path
        W2_ = O.random_weights("social", seed=21)
        pool2_ = make_pool("social", W2_)
        os.environ["TB2_SPARSE"] = "tc"
        ref2 = pool2_(*t_in).cpu().numpy()
        os.environ["TB2_SPARSE"] = args.mode
        out2 = pool2_(*t_in)
        torch.cuda.synchronize()
        d2 = np.abs(out2.cpu().numpy() - ref2)
        print("[%s] big ragged=%d two_layer pooled  max|mode - tc| = %.3e  mean %.3e  (ref max %.3f)  fuse2=%s" % (
            args.mode, ragged, float(np.nanmax(d2)), float(np.nanmean(d2)), float(np.nanmax(np.abs(ref2))),
            "on" if os.environ.get("TB2_FUSE2") == "1" else "off"), flush=True)
        if not ragged:
            # timing of the whole pool call (prepare + layer 1) with the cycle counters of the kernel
            os.environ["TB2_L1_DEBUG"] = "1"
            W2 = O.random_weights("social", seed=21)
            pool2 = make_pool("social", W2)
            for p_, name in ((pool, "one_layer(1024)"), (pool2, "two_layer(1024,256)")):
                for _ in range(5):
                    p_(*t_in)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    p_(*t_in)
                e1.record()
                torch.cuda.synchronize()
                print("[%s] %s pool call: %.1f us per call (prepare + MLP, %d calls)" % (
                    args.mode, name, 1e3 * e0.elapsed_time(e1) / args.iters, args.iters), flush=True)


if __name__ == "__main__":
    main()
