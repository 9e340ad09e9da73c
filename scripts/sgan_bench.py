"""S-GAN generator inference: k modes per call, 256 scenes x 20 pedestrians,
social pooling.  The encoder runs once, every mode decodes from a copy of its state.  One JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lstm_oracle as O
from oracle import sgan_oracle as SO
from trajnetplusplusbaselines_b200.lstm import GridBasedPooling
from trajnetplusplusbaselines_b200.sgan import SGAN, LSTMGenerator

kind = sys.argv[1] if len(sys.argv) > 1 else "social"
K = int(os.environ.get("TB2_SGAN_MODES", "3"))
B, N = 256, 20
Wg, _ = SO.sgan_weights(kind, 1)
spec = O.MODEL_SPECS[kind]
gen = LSTMGenerator(pool=GridBasedPooling(**spec) if spec else None)
sd = gen.state_dict()
sd.update({k: torch.from_numpy(v.copy()) for k, v in Wg.items()})
gen.load_state_dict(sd)
model = SGAN(generator=gen.cuda().eval(), k=K, d_steps=0).eval()
xy, bs = O.synthetic_scenes(B, N, seed=100)
obs = torch.from_numpy(xy[:9]).cuda()
bs_t = torch.from_numpy(bs)
goals = torch.zeros(xy.shape[1], 2)
with torch.no_grad():
    for _ in range(3):
        model(obs, goals, bs_t, n_predict=12)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    R = 10
    for _ in range(R):
        model(obs, goals, bs_t, n_predict=12)
    b.record()
    torch.cuda.synchronize()
ms = a.elapsed_time(b) / R
steps = 8 + K * 11
print(json.dumps({"workload": "S-GAN generator (%s pool), %d scenes x %d peds, k = %d modes" % (kind, B, N, K),
                  "ms_per_call": ms, "ped_steps_per_s": xy.shape[1] * steps / (ms * 1e-3),
                  "steps_per_call": steps, "reference_would_run_steps": K * 19}))
