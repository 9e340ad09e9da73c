"""Classical simulators (BASELINE configs[4]): 1 M pedestrians = 50 000 scenes x 20, 1000 steps,
sharded over 8 GPUs -> 6 250 scenes (125 000 pedestrians) per GPU.  Prints JSON lines with
ped-steps/s for social force (fp64) and ORCA (fp32) on this GPU's shard, plus the CPU restatement
timed on a bounded sample of the same workload."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from trajnetplusplusbaselines_b200.classical import orca, socialforce

scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 6250
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
N = 20
rng = np.random.RandomState(0)
A = scenes * N
offs = np.arange(0, A + 1, N)
pos = rng.randn(A, 2) * 2.0
ang = rng.rand(A) * 2 * np.pi
spd = 0.4 + rng.rand(A) * 1.2
vel = np.stack([spd * np.cos(ang), spd * np.sin(ang)], axis=1)
goal = pos + vel * 0.4 * 12 * 4
state = np.concatenate([pos, vel, goal], axis=1)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), out


st_dev = torch.as_tensor(state, dtype=torch.float64).cuda()
ms, out = timed(lambda: socialforce.simulate_batch(st_dev, offs.tolist(), n_steps=steps, sample_every=8))
print(json.dumps({"sim": "social_force_fp64", "scenes": scenes, "peds": A, "steps": steps, "ms": ms,
                  "ped_steps_per_s": A * steps / (ms * 1e-3), "finite": bool(torch.isfinite(out).all())}))
ms, out = timed(lambda: orca.simulate_batch(pos, vel, goal, spd, offs.tolist(), n_steps=steps, sample_every=8))
print(json.dumps({"sim": "orca_fp32", "scenes": scenes, "peds": A, "steps": steps, "ms": ms,
                  "ped_steps_per_s": A * steps / (ms * 1e-3), "finite": bool(torch.isfinite(out).all())}))

# CPU restatements on a bounded sample (single thread)
from oracle import classical_oracle as C
from oracle.build_c import orca_simulate
ns = 20
t0 = time.perf_counter()
for b in range(ns):
    C.sf_simulate(state[b * N:(b + 1) * N], n_steps=96)
t_sf = time.perf_counter() - t0
t0 = time.perf_counter()
for b in range(ns * 10):
    orca_simulate(pos[b * N:(b + 1) * N], vel[b * N:(b + 1) * N], goal[b * N:(b + 1) * N], spd[b * N:(b + 1) * N], n_steps=97)
t_orca = time.perf_counter() - t0
print(json.dumps({"cpu_restatement": {"sf_numpy_ped_steps_per_s": ns * N * 96 / t_sf,
                                      "orca_c_ped_steps_per_s": ns * 10 * N * 97 / t_orca, "cores": 1}}))

# BASELINE configs[0]: Kalman predictor, 64 scenes x 5 pedestrians, 9 observed -> 12 predicted; CPU only
# (host C++ EM + RTS smoother behind tb2_kalman_predict vs the numpy restatement, one thread each)
from trajnetplusplusbaselines_b200.classical import kalman
krng = np.random.RandomState(5)
tracks = []
for _ in range(64 * 5):
    p0, v = krng.randn(2) * 2.0, krng.randn(2) * 0.3
    tracks.append(p0 + np.arange(9)[:, None] * v + krng.randn(9, 2) * 0.05)
kalman.predict_tracks(tracks[:8], n_predict=12, n_samples=0)
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    pred = kalman.predict_tracks(tracks, n_predict=12, n_samples=0)
t_kf = (time.perf_counter() - t0) / reps
t0 = time.perf_counter()
for t in tracks[:32]:
    C.kalman_predict_track(t, n_predict=12, n_iter=10)
t_kf_np = (time.perf_counter() - t0) / 32 * len(tracks)
print(json.dumps({"sim": "kalman_host_cpp", "scenes": 64, "tracks": len(tracks), "ms": 1e3 * t_kf,
                  "ped_steps_per_s": len(tracks) * 12 / t_kf, "numpy_restatement_ped_steps_per_s": len(tracks) * 12 / t_kf_np,
                  "cores": 1, "finite": bool(np.isfinite(pred).all())}))
