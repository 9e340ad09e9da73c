"""D-LSTM training step (BASELINE configs[3]): teacher-forced forward + PredictionLoss + CUDA BPTT +
Adam, 256 scenes x 20 peds per GPU; with torchrun also one flat-bucket NCCL all-reduce per step.
Prints one JSON line (ped-steps/s = tracks x 19 / step time)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from oracle import lstm_oracle as O
from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling, PredictionLoss
from trajnetplusplusbaselines_b200.parallel import allreduce_gradients

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
kind = "directional"
B, N = 256, 20
W = O.random_weights(kind, seed=1)
model = LSTM(pool=GridBasedPooling(**O.MODEL_SPECS[kind]))
model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()})
model = model.cuda().train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)   # trainer.py:497 hyper-parameters
crit = PredictionLoss()
xy, bs = O.synthetic_scenes(B, N, seed=100 + rank)
scene = torch.from_numpy(xy).cuda()
bs_t = torch.from_numpy(bs)
targets = scene[9:21] - scene[8:20]
goals = torch.zeros(xy.shape[1], 2)


def step():
    rel, _ = model(scene[:9], goals, bs_t, scene[9:-1])
    loss = crit(rel[-12:], targets, bs_t) * B
    opt.zero_grad()
    loss.backward()
    if world > 1:
        allreduce_gradients(model.parameters())
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
K = 10
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(K):
    loss = step()
b.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    M = xy.shape[1]
    print(json.dumps({"workload": "D-LSTM train_batch (directional n=12 one_layer), %d scenes x %d peds per GPU" % (B, N),
                      "n_gpus": world, "ms_per_step": ms.item() / K, "wall_ms_per_step": 1e3 * wall / K,
                      "ped_steps_per_s": M * 19 * world * K / (ms.item() * 1e-3), "loss": float(loss.item())}))
if world > 1:
    dist.destroy_process_group()
