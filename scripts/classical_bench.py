"""Classical simulators, BASELINE configs[4]: 1 M pedestrians = 50 000 scenes x 20, 1000 steps, scenes sharded
over the GPUs of the launch (no collective on the data path; NCCL only for the barrier / max-time reduction).

    python scripts/classical_bench.py [total_scenes] [steps]                       # one GPU: its 1/8 shard by default
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/classical_bench.py 50000 1000   # the whole config

Prints one JSON line per simulator: whole-job ped-steps/s (device-timed, max over ranks), SM clocks sampled
during the timed region, and the per-ped-step work model; on one GPU also the CPU restatement on a bounded
sample.  Parity of both simulators is UNPINNED vs upstream (socialforce / rvo2 are not vendored): the numbers
say how fast this restatement of the published algorithms runs, nothing about upstream's results."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from bench import ClockSampler
from trajnetplusplusbaselines_b200.classical import orca, socialforce

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
device = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
total_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else (50000 if world > 1 else 6250)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
scenes = total_scenes // world + (1 if rank < total_scenes % world else 0)       # contiguous scene shard of this rank
N = 20
rng = np.random.RandomState(rank)
A = scenes * N
offs = np.arange(0, A + 1, N)
pos = rng.randn(A, 2) * 2.0
ang = rng.rand(A) * 2 * np.pi
spd = 0.4 + rng.rand(A) * 1.2
vel = np.stack([spd * np.cos(ang), spd * np.sin(ang)], axis=1)
goal = pos + vel * 0.4 * 12 * 4
state = np.concatenate([pos, vel, goal], axis=1)

# work per pedestrian-step at N = 20 (19 neighbours), counted from the kernels' source:
#   social force: 3 evaluations of the elliptic potential per neighbour (value + two forward differences), each 3 sqrt
#     + 1 exp + 1 div + ~22 fp64 mul/add, plus the field-of-view test (1 sqrt) -> ~75 fp64 FMA-class ops, 10 sqrt, 3 exp,
#     5 div per pair; x 19 pairs + ~40 for the integration.  State is on chip; HBM sees 16 B per sampled position.
#   ORCA: neighbour search over 19 agents (4 FFMA each + insertion), <= 10 ORCA lines (~60 fp32 ops, 2 sqrt, 2 div each),
#     linearProgram2/3 (data dependent, ~30-300 ops), 8 B per sampled position.
MODEL = {"social_force_fp64": {"fp64_ops": 19 * 75 + 40, "sqrt": 19 * 10 + 3, "exp": 19 * 3, "div": 19 * 5 + 3,
                               "hbm_bytes": 16.0 / 8},
         "orca_fp32": {"fp32_ops": 19 * 6 + 10 * 60 + 150, "sqrt": 25, "div": 25, "hbm_bytes": 8.0 / 8}}


def timed(fn):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.2)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item(), out, clocks


def report(name, ms, out, clocks):
    fin = torch.tensor([int(bool(torch.isfinite(out).all()))], device=device)
    if world > 1:
        dist.all_reduce(fin, op=dist.ReduceOp.MIN)
    if rank == 0:
        peds = total_scenes * N
        v = peds * steps / (ms * 1e-3)
        line = {"sim": name, "n_gpus": world, "scenes": total_scenes, "peds": peds, "steps": steps, "ms": ms,
                "ped_steps_per_s": v, "finite": bool(fin.item()), "clocks": clocks, "work_per_ped_step": MODEL[name],
                "parity": "unpinned vs upstream (restatement of the published algorithm)"}
        ops = MODEL[name].get("fp64_ops") or MODEL[name].get("fp32_ops")
        line["achieved_Tops"] = v * ops / 1e12
        line["hbm_GBps"] = v * MODEL[name]["hbm_bytes"] / 1e9
        print(json.dumps(line), flush=True)


st_dev = torch.as_tensor(state, dtype=torch.float64).to(device)
ms, out, clocks = timed(lambda: socialforce.simulate_batch(st_dev, offs.tolist(), n_steps=steps, sample_every=8, device=device))
report("social_force_fp64", ms, out, clocks)
ms, out, clocks = timed(lambda: orca.simulate_batch(pos, vel, goal, spd, offs.tolist(), n_steps=steps, sample_every=8, device=device))
report("orca_fp32", ms, out, clocks)

if world == 1:
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
                                           "orca_c_ped_steps_per_s": ns * 10 * N * 97 / t_orca, "cores": 1,
                                           "sample": "%d / %d scenes x 20 peds x ~96 steps" % (ns, ns * 10)}}))
if world > 1:
    dist.destroy_process_group()
