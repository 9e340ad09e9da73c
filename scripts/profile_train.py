"""Where does a D-LSTM training step go?  Phase timings (host wall with a sync after each phase)
and the library's per-kernel CUDA-event breakdown for forward+backward."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lstm_oracle as O
from trajnetplusplusbaselines_b200 import _lib
from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling, PredictionLoss

kind = sys.argv[1] if len(sys.argv) > 1 else "directional"
B, N = int(os.environ.get("TB2_BENCH_SCENES", "256")), 20
W = O.random_weights(kind, seed=1)
model = LSTM(pool=GridBasedPooling(**O.MODEL_SPECS[kind]) if kind != "vanilla" else None)
model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()})
model = model.cuda().train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
crit = PredictionLoss()
xy, bs = O.synthetic_scenes(B, N, seed=100)
scene = torch.from_numpy(xy).cuda()
bs_t = torch.from_numpy(bs)
targets = scene[9:21] - scene[8:20]
goals = torch.zeros(xy.shape[1], 2)
phases = {}


def tick(name, t0):
    torch.cuda.synchronize()
    phases[name] = phases.get(name, 0.0) + (time.perf_counter() - t0)
    return time.perf_counter()


def step(timed):
    t = time.perf_counter()
    rel, _ = model(scene[:9], goals, bs_t, scene[9:-1])
    if timed:
        t = tick("forward", t)
    loss = crit(rel[-12:], targets, bs_t) * B
    if timed:
        t = tick("loss", t)
    opt.zero_grad()
    loss.backward()
    if timed:
        t = tick("backward", t)
    opt.step()
    if timed:
        t = tick("adam", t)


for _ in range(3):
    step(False)
torch.cuda.synchronize()
K = 10
for _ in range(K):
    step(True)
print({k: round(1e3 * v / K, 3) for k, v in phases.items()}, "ms per step (phase-synchronised)")
lib = _lib.load()
lib.tb2_profile_begin()
step(False)
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 16)
_lib.check(lib.tb2_profile_end(buf, len(buf)))
import json
prof = json.loads(buf.value.decode())
tot = 0.0
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"]):
    print("%-28s %4d launches %8.1f us total" % (k, v["launches"], 1e3 * v["total_ms"]))
    tot += v["total_ms"]
print("library kernels total %.3f ms" % tot)
