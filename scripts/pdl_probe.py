"""Which part of the host-tensor path interacts badly with programmatic dependent launch?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lstm_oracle as O
from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling

W = O.random_weights("social", seed=1)
model = LSTM(pool=GridBasedPooling(**O.MODEL_SPECS["social"]))
model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()})
model = model.cuda().eval()
xy, bs = O.synthetic_scenes(256, 20, seed=100)
obs_host = torch.from_numpy(xy[:9]).float().pin_memory()
obs_dev = obs_host.cuda()
bs_t = torch.from_numpy(bs)
goals = torch.zeros(xy.shape[1], 2)
pin_out = torch.empty((19, xy.shape[1], 2), dtype=torch.float32).pin_memory()


import gc
GC_OFF = os.environ.get("PROBE_GC_OFF") == "1"


def timeit(name, fn, n=12):
    for _ in range(3):
        fn()
    if GC_OFF:
        gc.collect()
        gc.disable()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    gc.enable()
    print("%-52s %s" % (name, " ".join("%.1f" % t for t in ts)))


with torch.no_grad():
    timeit("A resident in/out, sync per forward", lambda: model(obs_dev, goals, bs_t, n_predict=12))
    timeit("B H2D(pinned) in, device out", lambda: model(obs_host.cuda(non_blocking=True), goals, bs_t, n_predict=12))

    def c():
        rel, pred = model(obs_dev, goals, bs_t, n_predict=12)
        pin_out.copy_(pred, non_blocking=True)
    timeit("C resident in, D2H(pinned) out", c)
    timeit("D full host path (LSTM.forward with host tensors)", lambda: model(obs_host, goals, bs_t, n_predict=12))
    obs_pageable = torch.from_numpy(xy[:9]).float()
    timeit("E pageable host input (staging copy), host out", lambda: model(obs_pageable, goals, bs_t, n_predict=12))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def f():
        flush.zero_()
        torch.cuda.synchronize()
        return model(obs_pageable, goals, bs_t, n_predict=12)
    timeit("F as bench.py: L2 flush + sync, then E", f)

    def g():
        flush.zero_()
        return model(obs_dev, goals, bs_t, n_predict=12)
    timeit("G L2 flush (no sync) then resident forward", g)
