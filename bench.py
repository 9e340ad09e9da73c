#!/usr/bin/env python
"""Benchmark of the TrajNet++ hot path on B200 (driver contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle port)

A "step" is one pass of the hot path over one batch of synthetic scenes: one call of
LSTM.forward = (obs-1) + (pred-1) = 19 recurrence steps for every track of the batch.
Metric: pedestrian-steps / second = tracks x 19 x K / time.
Workload at every N: BASELINE.json configs[2] -- Social-LSTM (--type social --n 16
--embedding_arch two_layer --layer_dims 1024), 256 scenes x 20 pedestrians PER GPU (weak
scaling: scenes are independent, no data-path collective), T = 9 observed + 12 predicted.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KIND = "social"
SCENES_PER_GPU = 256
PEDS = 20
OBS, PRED = 9, 12
STEPS_PER_FORWARD = (OBS - 1) + (PRED - 1)       # 19
STATE_BYTES_PER_PED_STEP = 2092                   # xy 16 + h,c in 1024 + h,c out 1024 + normal 20 + pos 8
DENSE_FLOP_PER_PED_STEP = {                       # SURVEY.md 8d, dense-equivalent forward FLOPs
    "sparse_layer1": 2 * 4096 * 1024,             # first Linear of the grid embedding (4096 -> 1024)
    "sparse_layer1_mma": 2 * 4096 * 1024,
    "sparse_layer1_tc": 2 * 4096 * 1024,
    "sparse_layer1_pair": 2 * 4096 * 1024,
    "sparse_layer1_solo": 2 * 4096 * 1024,
    "dense_layer": 2 * 1024 * 256,
    "dense_layer_tc": 2 * 1024 * 256,
    "lstm_gates": 2 * (64 + 256 + 128) * 512 + 2 * 128 * 5,
    "lstm_gates_tc": 2 * (64 + 256 + 128) * 512 + 2 * 128 * 5,
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.proc = None
        self.lines = []

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)),
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(rank, scenes, seed=0):
    from oracle import lstm_oracle as O      # synthetic generator only (shared with the tests)
    xy, bs = O.synthetic_scenes(scenes, PEDS, n_frames=OBS + PRED, seed=seed + 1000 * rank)
    return xy, bs


def cpu_oracle_run(scenes, threads_note=True):
    """The oracle port (numpy fp32, BLAS threads = all host cores) on `scenes` scenes."""
    from oracle import lstm_oracle as O
    W = O.random_weights(KIND, seed=1)
    xy, bs = make_inputs(0, scenes)
    cfg = O.pool_config(KIND)
    t0 = time.perf_counter()
    O.forward(W, cfg, xy[:OBS], bs, n_predict=PRED)
    dt = time.perf_counter() - t0
    return scenes * PEDS * STEPS_PER_FORWARD / dt, dt


def run_reference(args):
    """--impl reference: the CPU restatement of the reference (oracle port; the Python reference
    itself cannot travel to the GPU box) on the host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    sample_scenes = 64
    for _ in range(args.warmup):
        cpu_oracle_run(8)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_oracle_run(sample_scenes)
        t_total += dt
    value = sample_scenes * PEDS * STEPS_PER_FORWARD * args.steps / t_total
    line = {
        "impl": "reference", "metric": "pedestrian-steps/sec", "value": value, "unit": "ped-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Social-LSTM inference (BASELINE configs[2]): type=social n=16 two_layer 1024, "
                               "N=20, T=9+12, free-running; CPU sample = %d scenes per step" % sample_scenes},
        "cpu_baseline": {"value": value, "unit": "ped-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d scenes x 20 peds x 19 steps per step, numpy fp32 oracle, BLAS threads=all" % sample_scenes},
        "e2e": {"value": value, "unit": "ped-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenes", type=int, default=SCENES_PER_GPU, help="scenes per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from oracle import lstm_oracle as O
    from trajnetplusplusbaselines_b200 import _lib
    from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    # model: random-init weights of the BASELINE architecture (seeded, same as the CPU arm)
    W = O.random_weights(KIND, seed=1)
    model = LSTM(pool=GridBasedPooling(**O.MODEL_SPECS[KIND]))
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()})
    model = model.to(device).eval()

    xy, bs = make_inputs(rank, args.scenes)
    M = xy.shape[1]
    observed_host = torch.from_numpy(xy[:OBS]).pin_memory()
    observed_dev = observed_host.to(device)
    goals = torch.zeros(M, 2)
    bs_t = torch.from_numpy(bs)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)     # > 126 MB L2

    def step_resident():
        with torch.no_grad():
            return model(observed_dev, goals, bs_t, n_predict=PRED)

    def step_e2e():
        with torch.no_grad():
            return model(observed_host, goals, bs_t, n_predict=PRED)      # H2D in, D2H out inside

    # ---- device-resident arm ---------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = int(lib.tb2_launch_count())
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (untimed)
        starts[i].record()
        step_resident()
        stops[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = int(lib.tb2_launch_count()) - launches0
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    clocks = sampler.stop()

    # ---- end-to-end arm (host buffers, copies inside the timed region) -----------------------
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    def e2e_pass():
        per_step = []
        for i in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            out = step_e2e()               # returns host tensors after a stream sync
            per_step.append(1e3 * (time.perf_counter() - t0))
        return per_step, out

    import gc
    gc.collect()
    per_step, (rel, pred) = e2e_pass()
    e2e_remeasured = False
    if max(per_step) > 5.0 * sorted(per_step)[len(per_step) // 2]:
        # a host-side stall (another tenant on the box, a descheduled process) hit one wall-clock
        # timed step: the whole pass is measured again, once, and that second pass is what counts
        per_step, (rel, pred) = e2e_pass()
        e2e_remeasured = True
    e2e_ms = sum(per_step)
    barrier()
    h2d = observed_host.numel() * 4 + bs_t.numel() * 8
    d2h = (rel.numel() + pred.numel()) * 4

    # ---- per-kernel CUDA-event timing for the roofline (separate pass, events on the launch stream)
    prof_iters = 3
    lib.tb2_profile_begin()
    for _ in range(prof_iters):
        flush.zero_()
        step_resident()
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(lib.tb2_profile_end(buf, len(buf)))
    prof = json.loads(buf.value.decode())

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = t.tolist()

    if rank == 0:
        ped_steps = M * STEPS_PER_FORWARD * world          # every rank runs the same shape
        value = ped_steps * args.steps / (ms_max * 1e-3)
        e2e_value = ped_steps * args.steps / (e2e_ms_max * 1e-3)
        hbm, tf_burst, tf_sust, how = peaks()
        total_ms = sum(v["total_ms"] for v in prof.values()) or 1.0
        dom = max(prof, key=lambda k: prof[k]["total_ms"])
        dom_avg_ms = prof[dom]["total_ms"] / prof[dom]["launches"]
        kern = {}
        for name, v in prof.items():
            avg = v["total_ms"] / v["launches"]
            kern[name] = {"avg_us": 1e3 * avg, "launches_per_forward": v["launches"] / prof_iters,
                          "share": v["total_ms"] / total_ms}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "round1_traffic.json")
        if os.path.exists(tpath):      # dram bytes per launch from the committed ncu --set full capture
            traffic = json.load(open(tpath))["bytes_per_launch"].get(dom)
        if dom in DENSE_FLOP_PER_PED_STEP:
            flops = DENSE_FLOP_PER_PED_STEP[dom] * M
            achieved = flops / (dom_avg_ms * 1e-3) / 1e12
            roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": tf_sust,
                        "unit": "TFLOP/s", "frac": achieved / tf_sust, "traffic": traffic,
                        "peak_source": how + " bf16 sustained (kernel timed inside a long step)",
                        "note": "achieved = dense algorithmic FLOPs of the 4096->1024 grid Linear (SURVEY 8d: "
                                "2*4096*1024 per ped-step) / CUDA-event time; the kernel issues 3 bf16 passes "
                                "(hi/lo split for the 1e-4 m gate), so 1/3 of peak is its ceiling"}
        else:
            bytes_ = STATE_BYTES_PER_PED_STEP * M
            achieved = bytes_ / (dom_avg_ms * 1e-3) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                        "frac": achieved / hbm, "traffic": traffic, "peak_source": how}
        # state-streaming view of the whole step (all kernels of one recurrence step)
        step_ms = total_ms / prof_iters / STEPS_PER_FORWARD
        roofline["step_hbm"] = {"achieved": STATE_BYTES_PER_PED_STEP * M / (step_ms * 1e-3) / 1e9,
                                "peak": hbm, "unit": "GB/s",
                                "frac": STATE_BYTES_PER_PED_STEP * M / (step_ms * 1e-3) / 1e9 / hbm}
        roofline["kernels"] = kern
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_oracle_run(8)
            v, dt = cpu_oracle_run(SCENES_PER_GPU if args.scenes >= SCENES_PER_GPU else args.scenes)
            cpu = {"value": v, "unit": "ped-steps/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": "one forward of the same workload (%d scenes x 20 peds x 19 steps, %.1f s), "
                             "numpy fp32 oracle, BLAS threads=all" % (min(args.scenes, SCENES_PER_GPU), dt)}
        line = {
            "metric": "pedestrian-steps/sec", "value": value, "unit": "ped-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "Social-LSTM inference (BASELINE configs[2]): type=social n=16 cell_side=0.6 "
                                   "two_layer 1024 -> 256, latent 16, hidden 128",
                       "scenes_per_gpu": args.scenes, "peds_per_scene": PEDS, "obs": OBS, "pred": PRED,
                       "recurrence_steps_per_step": STEPS_PER_FORWARD, "parallelism": "scenes sharded x%d, no collective" % world,
                       "l2": "256 MiB memset between timed iterations (untimed); inputs are smaller than L2"},
            "e2e": {"value": e2e_value, "unit": "ped-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_max / args.steps, "remeasured_after_host_stall": e2e_remeasured},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
