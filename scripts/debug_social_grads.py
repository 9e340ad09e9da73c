"""Per-parameter deviation of the CUDA social backward from the reference goldens."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from oracle import lstm_oracle as O
from oracle.make_train_golden import TRAIN_CASES, N_SAMPLES
from test_training import _cuda_train_step

G = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "train_golden.npz"))
for case in TRAIN_CASES:
    name, kind, B, N, ragged, nan_tracks, dseed, wseed = case
    if "social" not in kind:
        continue
    xy, bs = O.synthetic_scenes(B, N, seed=dseed, ragged=ragged, nan_tracks=nan_tracks)
    W = O.random_weights(kind, seed=wseed)
    model, loss = _cuda_train_step(kind, W, xy, bs)
    print(name, "loss", loss, "ref", float(G[name + "/loss"][0]))
    for pname, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.cpu().numpy()
        key = name + "/" + pname
        if key + "/full" in G:
            ref = G[key + "/full"]
            dev = np.abs(g - ref).max()
            scale = np.abs(ref).max()
        else:
            idx = np.random.RandomState(12345).randint(0, g.size, size=N_SAMPLES)
            ref = G[key + "/samples"]
            dev = np.abs(g.reshape(-1)[idx] - ref).max()
            scale = np.abs(ref).max()
        print("   %-50s max|ref| %.3e  max dev %.3e  rel %.2e" % (pname, scale, dev, dev / max(scale, 1e-12)))
