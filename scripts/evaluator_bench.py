"""Batched evaluator path: ndjson-style scenes -> LSTMPredictor -> ndjson.
Per-scene calls (what lstm/trajnet_evaluator.py:18,61 does per joblib worker) vs one predict_batch
call for all scenes.  Prints one JSON line (scenes/s, end to end incl. paths_to_xy and the writer)."""
import json
import os
import sys
import tempfile
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import lstm_oracle as O
from trajnetplusplusbaselines_b200.data import TrackRow, write_predictions
from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling, LSTMPredictor

kind = sys.argv[1] if len(sys.argv) > 1 else "social"
B, N = int(os.environ.get("TB2_BENCH_SCENES", "1024")), 20
xy, bs = O.synthetic_scenes(B, N, seed=11)
scenes = []
for b in range(B):
    paths = []
    for p in range(bs[b], bs[b + 1]):
        paths.append([TrackRow(10 * t, 1000 * b + int(p - bs[b]), float(xy[t, p, 0]), float(xy[t, p, 1]))
                      for t in range(9) if not np.isnan(xy[t, p, 0])])
    scenes.append(("synthetic", b, [p for p in paths if p]))
W = O.random_weights(kind, seed=1)
model = LSTM(pool=GridBasedPooling(**O.MODEL_SPECS[kind]) if O.MODEL_SPECS[kind] else None)
model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()})
predictor = LSTMPredictor(model.cuda())
args = types.SimpleNamespace(normalize_scene=False)
n_single = min(B, 128)
for _, _, paths in scenes[:8]:
    predictor(paths, np.zeros((len(paths), 2)), n_predict=12, obs_length=9, args=args)
torch.cuda.synchronize()
t0 = time.perf_counter()
single = [predictor(paths, np.zeros((len(paths), 2)), n_predict=12, obs_length=9, args=args)
          for _, _, paths in scenes[:n_single]]
torch.cuda.synchronize()
t_single = (time.perf_counter() - t0) / n_single
predictor.predict_batch([paths for _, _, paths in scenes[:64]], n_predict=12, obs_length=9, args=args)
torch.cuda.synchronize()
t0 = time.perf_counter()
batched = predictor.predict_batch([paths for _, _, paths in scenes], n_predict=12, obs_length=9, args=args)
torch.cuda.synchronize()
t_batch = time.perf_counter() - t0
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    write_predictions(batched, scenes, os.path.join(d, "pred.ndjson"))
    t_write = time.perf_counter() - t0
# file -> file through evaluate_file (column pipeline: native ndjson parse, predict_batch_xy, native formatting)
from trajnetplusplusbaselines_b200.data import SceneRow, trajnet_line
from trajnetplusplusbaselines_b200.evaluator import evaluate_file
with tempfile.TemporaryDirectory() as d:
    infile, outfile = os.path.join(d, "in.ndjson"), os.path.join(d, "out.ndjson")
    with open(infile, "w") as f:
        for _, b, paths in scenes:
            f.write(trajnet_line(SceneRow(b, paths[0][0].pedestrian, 20000 * b, 20000 * b + 200, 2.5, 0)) + "\n")
            for path in paths:
                for r in path:
                    f.write(trajnet_line(TrackRow(20000 * b + r.frame, r.pedestrian, r.x, r.y)) + "\n")
    evaluate_file(predictor, infile, outfile, args=args)          # warm-up (layout / pinned pools)
    t0 = time.perf_counter()
    n_file = evaluate_file(predictor, infile, outfile, args=args)
    torch.cuda.synchronize()
    t_file = time.perf_counter() - t0
dev = max(float(np.abs(single[i][0][0] - batched[i][0][0]).max()) for i in range(n_single))
print(json.dumps({"workload": "%s evaluator path, %d scenes x %d peds, obs 9 -> pred 12" % (kind, B, N),
                  "per_scene_call_scenes_per_s": 1.0 / t_single, "predict_batch_scenes_per_s": B / t_batch,
                  "speedup": t_single * B / t_batch, "ndjson_write_scenes_per_s": B / t_write,
                  "evaluate_file_scenes_per_s": n_file / t_file,
                  "max_abs_dev_single_vs_batched_m": dev}))
