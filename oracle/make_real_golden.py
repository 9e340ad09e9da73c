"""Real-scene fixtures (SURVEY.md 8d variant v4): scenes of the reference's shipped
DATA_BLOCK/trajdata/train/*.ndjson after drop_distant (lstm/lstm.py:16-22), run through the
UNMODIFIED reference.  Inputs (a few scenes, float32) and reference outputs are stored in
tests/golden/real_scenes.npz.   python -m oracle.make_real_golden      TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import lstm_oracle as O                     # noqa: E402
from oracle.ref_shim import import_reference, REFERENCE_ROOT   # noqa: E402
from oracle.make_golden import build_reference_model     # noqa: E402
from trajnetplusplusbaselines_b200.data import paths_to_xy, read_ndjson_scenes   # noqa: E402

FILES = ["biwi_hotel.ndjson", "crowds_students001.ndjson"]
SCENES_PER_FILE = 6
KINDS = [("directional", 21), ("social_small", 22), ("vanilla", 23), ("social", 24), ("occupancy", 25)]


def load_scenes():
    import_reference()
    from trajnetbaselines.lstm.lstm import drop_distant
    xys = []
    for fn in FILES:
        path = os.path.join(REFERENCE_ROOT, "DATA_BLOCK", "trajdata", "train", fn)
        picked = 0
        for i, (scene_id, paths) in enumerate(read_ndjson_scenes(path)):
            if i % 97 != 0:
                continue
            xy = paths_to_xy(paths)
            if xy.shape[0] != 21:
                continue
            xy, _ = drop_distant(xy)
            xys.append(xy.astype(np.float32))
            picked += 1
            if picked == SCENES_PER_FILE:
                break
    return xys


def main():
    xys = load_scenes()
    sizes = [x.shape[1] for x in xys]
    bs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    xy = np.concatenate(xys, axis=1)
    out = {"xy": xy, "batch_split": bs}
    print("scenes:", len(xys), "sizes:", sizes, "NaN fraction: %.2f" % np.isnan(xy[:, :, 0]).mean())
    for kind, wseed in KINDS:
        W = O.random_weights(kind, seed=wseed)
        model = build_reference_model(kind, W)
        with torch.no_grad():
            rel, pred = model(torch.from_numpy(xy[:9]), torch.zeros(xy.shape[1], 2), torch.from_numpy(bs), n_predict=12)
            rel_t, pred_t = model(torch.from_numpy(xy[:9]), torch.zeros(xy.shape[1], 2), torch.from_numpy(bs),
                                  prediction_truth=torch.from_numpy(xy[9:20]).clone())
        out[kind + "/pred_free"] = pred.numpy()
        out[kind + "/rel_free"] = rel.numpy()
        out[kind + "/pred_teacher"] = pred_t.numpy()
        # oracle agreement printed for information
        _, pred_o = O.forward(W, O.pool_config(kind), xy[:9], bs, n_predict=12)
        print(kind, "oracle vs reference max |d pos| = %.2e" % np.nanmax(np.abs(pred_o - pred.numpy())))
    path = os.path.join(ROOT, "tests", "golden", "real_scenes.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
