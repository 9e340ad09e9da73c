"""Host I/O of the batched evaluator path, no GPU needed: the row pipeline
(read_ndjson_scenes -> preprocess_test -> paths_to_xy, write_predictions) against the column pipeline
(load_test_scenes_xy, write_predictions_xy: native text passes of csrc/ndjson.cu) on one synthetic DATA_BLOCK-style
file.  Prints one JSON line (scenes/s, best of 3)."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trajnetplusplusbaselines_b200.data import (SceneRow, TrackRow, load_test_scenes_xy, paths_to_xy, preprocess_test,
                                                read_ndjson_scenes, trajnet_line, write_predictions, write_predictions_xy)

B, N = int(os.environ.get("TB2_BENCH_SCENES", "1024")), 20
rng = np.random.RandomState(0)
d = tempfile.mkdtemp()
fn = os.path.join(d, "in.ndjson")
with open(fn, "w") as f:
    for sid in range(B):
        start, vel = rng.randn(N, 2) * 3, rng.randn(N, 2) * 0.2
        frames = [1000 * sid + 10 * t for t in range(21)]
        f.write(trajnet_line(SceneRow(sid, 100 * sid, frames[0], frames[-1], 2.5, 0)) + "\n")
        for p in range(N):
            for t, fr in enumerate(frames):
                f.write(trajnet_line(TrackRow(fr, 100 * sid + p, start[p, 0] + vel[p, 0] * t, start[p, 1] + vel[p, 1] * t)) + "\n")


def best(fn_, reps=3):
    out, t = None, float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn_()
        t = min(t, time.perf_counter() - t0)
    return out, t


def rows_in():
    scenes = [("f", sid, preprocess_test(paths, 9)) for sid, paths in read_ndjson_scenes(fn)]
    return scenes, [paths_to_xy(p) for _, _, p in scenes]


(rows, xys), t_rows_in = best(rows_in)
cols, t_cols_in = best(lambda: load_test_scenes_xy(fn, 9))
assert all(np.array_equal(a, b[0], equal_nan=True) for a, b in zip(xys, cols))
preds = [{0: [rng.randn(12, 2) * 5, rng.randn(12, N - 1, 2) * 5]} for _ in range(B)]
k = [0]


def out_name():
    k[0] += 1
    return os.path.join(d, "out%d.ndjson" % k[0])


_, t_rows_out = best(lambda: write_predictions(preds, rows, out_name()))
_, t_cols_out = best(lambda: write_predictions_xy(preds, [m for _, m in cols], out_name()))
a, b = out_name(), out_name()
write_predictions(preds, rows, a)
write_predictions_xy(preds, [m for _, m in cols], b)
assert open(a, "rb").read() == open(b, "rb").read()
print(json.dumps({"workload": "%d scenes x %d peds x 21 frames, %.1f MB in, %.1f MB out" % (B, N, os.path.getsize(fn) / 1e6, os.path.getsize(a) / 1e6),
                  "host": "build container CPU (%d cores), no GPU involved" % os.cpu_count(),
                  "read_rows_scenes_per_s": B / t_rows_in, "read_columns_scenes_per_s": B / t_cols_in,
                  "write_rows_scenes_per_s": B / t_rows_out, "write_columns_scenes_per_s": B / t_cols_out,
                  "host_io_total_rows_s": t_rows_in + t_rows_out, "host_io_total_columns_s": t_cols_in + t_cols_out,
                  "outputs_byte_identical": True}))
