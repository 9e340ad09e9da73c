"""Where does the host-side time of one end-to-end LSTM.forward go? (run on the GPU box)"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lstm_oracle as O
from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling

kind = "social"
W = O.random_weights(kind, seed=1)
model = LSTM(pool=GridBasedPooling(**O.MODEL_SPECS[kind]))
model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()})
model = model.cuda().eval()
xy, bs = O.synthetic_scenes(256, 20, seed=0)
obs_host = torch.from_numpy(xy[:9]).pin_memory()
obs_dev = obs_host.cuda()
goals = torch.zeros(xy.shape[1], 2)
bs_t = torch.from_numpy(bs)
with torch.no_grad():
    for _ in range(3):
        model(obs_host, goals, bs_t, n_predict=12)
    torch.cuda.synchronize()
    for name, obs in (("resident", obs_dev), ("host", obs_host)):
        t0 = time.perf_counter()
        for _ in range(5):
            out = model(obs, goals, bs_t, n_predict=12)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("%s: host-side %.3f ms / forward, +%.3f ms to drain" % (name, 1e3 * (t1 - t0) / 5, 1e3 * (t2 - t1)))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        model(obs_host, goals, bs_t, n_predict=12)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

# ---- bench-like loops: what does the L2 flush / the sync before each step cost? -------------
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(label, obs, do_flush, sync_before, n=10):
    with torch.no_grad():
        for _ in range(2):
            model(obs, goals, bs_t, n_predict=12)
        torch.cuda.synchronize()
        tot = 0.0
        ev = []
        for _ in range(n):
            if do_flush:
                flush.zero_()
            if sync_before:
                torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            model(obs, goals, bs_t, n_predict=12)
            b.record()
            torch.cuda.synchronize()
            tot += time.perf_counter() - t0
            ev.append(a.elapsed_time(b))
        print("%-34s wall %.3f ms  gpu(events) %.3f ms" % (label, 1e3 * tot / n, sum(ev) / n))


timed("resident  noflush nosync", obs_dev, False, False)
timed("resident  flush   nosync", obs_dev, True, False)
timed("resident  flush   sync", obs_dev, True, True)
timed("host      noflush sync", obs_host, False, True)
timed("host      flush   sync", obs_host, True, True)
