import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lstm_oracle as O
from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling, PredictionLoss
for fused in (False, True):
    kind = "directional"
    W = O.random_weights(kind, seed=1)
    model = LSTM(pool=GridBasedPooling(**O.MODEL_SPECS[kind]))
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()})
    model = model.cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=fused)
    xy, bs = O.synthetic_scenes(64, 20, seed=100)
    scene = torch.from_numpy(xy).cuda(); bs_t = torch.from_numpy(bs)
    targets = scene[9:21] - scene[8:20]; goals = torch.zeros(xy.shape[1], 2)
    crit = PredictionLoss(); losses = []
    p0 = next(model.parameters())
    for it in range(6):
        rel, _ = model(scene[:9], goals, bs_t, scene[9:-1])
        loss = crit(rel[-12:], targets, bs_t) * 64
        opt.zero_grad(); loss.backward()
        v0 = p0._version
        opt.step()
        losses.append(round(loss.item(), 4))
    print("fused", fused, "version bump per step:", p0._version - v0, "losses", losses)
