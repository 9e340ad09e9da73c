"""Inference throughput of every BASELINE LSTM config on one GPU (256 scenes x 20 peds, T = 9 + 12,
free-running; `python scripts/configs_bench.py <scenes>` for another batch), device-resident inputs, plus the
per-kernel CUDA-event breakdown and the state-streaming roofline of the step (2092 B per ped-step)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lstm_oracle as O
from trajnetplusplusbaselines_b200 import _lib
from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling

lib = _lib.load()
SCENES = int(sys.argv[1]) if len(sys.argv) > 1 else 256
HBM_GBS = 6592.6
try:
    HBM_GBS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
xy, bs = O.synthetic_scenes(SCENES, 20, seed=0)
obs = torch.from_numpy(xy[:9]).cuda()
goals = torch.zeros(xy.shape[1], 2)
bs_t = torch.from_numpy(bs)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for kind in ("vanilla", "occupancy", "directional", "social"):
    spec = O.MODEL_SPECS[kind]
    model = LSTM(pool=GridBasedPooling(**spec) if spec else None)
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in O.random_weights(kind, seed=1).items()})
    model = model.cuda().eval()
    with torch.no_grad():
        for _ in range(3):
            model(obs, goals, bs_t, n_predict=12)
        torch.cuda.synchronize()
        ms = 0.0
        K = 10
        for _ in range(K):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            model(obs, goals, bs_t, n_predict=12)
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        lib.tb2_profile_begin()
        model(obs, goals, bs_t, n_predict=12)
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(lib.tb2_profile_end(buf, len(buf)))
        prof = json.loads(buf.value.decode())
    step_s = ms / K * 1e-3 / 19
    hbm = 2092.0 * xy.shape[1] / step_s / 1e9
    print(json.dumps({"config": kind, "scenes": SCENES, "ms_per_forward": ms / K, "ped_steps_per_s": xy.shape[1] * 19 * K / (ms * 1e-3),
                      "step_hbm": {"achieved_GBps": round(hbm, 1), "peak_GBps": HBM_GBS, "frac": round(hbm / HBM_GBS, 4),
                                   "note": "state bytes of one recurrence step (2092 B x tracks) / step time"},
                      "kernels_us": {k: round(1e3 * v["total_ms"] / v["launches"], 1) for k, v in prof.items()}}))
