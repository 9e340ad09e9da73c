"""Social-LSTM training step (forward + PredictionLoss + BPTT through the hidden-state scatter +
Adam), 256 scenes x 20 pedestrians on one GPU.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lstm_oracle as O
from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling, PredictionLoss

B, N = int(os.environ.get("TB2_BENCH_SCENES", "256")), 20
W = O.random_weights("social", seed=1)
model = LSTM(pool=GridBasedPooling(**O.MODEL_SPECS["social"]))
model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()})
model = model.cuda().train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
crit = PredictionLoss()
xy, bs = O.synthetic_scenes(B, N, seed=100)
scene = torch.from_numpy(xy).cuda()
bs_t = torch.from_numpy(bs)
targets = scene[9:21] - scene[8:20]
goals = torch.zeros(xy.shape[1], 2)


def step():
    rel, _ = model(scene[:9], goals, bs_t, scene[9:-1])
    loss = crit(rel[-12:], targets, bs_t) * B
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
K = 5
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(K):
    loss = step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
print(json.dumps({"workload": "Social-LSTM train_batch (social n=16 two_layer 1024), %d scenes x %d peds" % (B, N),
                  "n_gpus": 1, "ms_per_step": ms, "ped_steps_per_s": xy.shape[1] * 19 / (ms * 1e-3),
                  "loss": float(loss.item()), "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}))
