"""Real scenes shipped with the reference (DATA_BLOCK/trajdata/train, after drop_distant): ragged
scene sizes 7..56, tracks entering / leaving mid-sequence.  Reference outputs are committed in
tests/golden/real_scenes.npz (oracle/make_real_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import lstm_oracle as O
from oracle.make_real_golden import KINDS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def real():
    return np.load(os.path.join(ROOT, "tests", "golden", "real_scenes.npz"))


@pytest.mark.parametrize("kind,wseed", KINDS)
def test_oracle_on_real_scenes(real, kind, wseed):
    xy, bs = real["xy"], real["batch_split"]
    W = O.random_weights(kind, seed=wseed)
    _, pred = O.forward(W, O.pool_config(kind), xy[:9], bs, n_predict=12)
    _, pred_t = O.forward(W, O.pool_config(kind), xy[:9], bs, prediction_truth=xy[9:20])
    for got, key in ((pred, "/pred_free"), (pred_t, "/pred_teacher")):
        ref = real[kind + key]
        assert (np.isnan(got) == np.isnan(ref)).all()
        assert np.nanmax(np.abs(got - ref)) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("kind,wseed", KINDS)
def test_cuda_on_real_scenes(real, kind, wseed):
    from trajnetplusplusbaselines_b200.lstm import LSTM, GridBasedPooling
    xy, bs = real["xy"], real["batch_split"]
    W = O.random_weights(kind, seed=wseed)
    spec = O.MODEL_SPECS[kind]
    model = LSTM(pool=GridBasedPooling(**spec) if spec is not None else None)
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()})
    model = model.cuda().eval()
    M = xy.shape[1]
    with torch.no_grad():
        rel, pred = model(torch.from_numpy(xy[:9]), torch.zeros(M, 2), torch.from_numpy(bs), n_predict=12)
        _, pred_t = model(torch.from_numpy(xy[:9]), torch.zeros(M, 2), torch.from_numpy(bs),
                          prediction_truth=torch.from_numpy(xy[9:20]).clone())
    for got, key in ((pred, "/pred_free"), (pred_t, "/pred_teacher"), (rel, "/rel_free")):
        ref = real[kind + key]
        got = got.numpy()
        assert (np.isnan(got) == np.isnan(ref)).all()
        d = np.abs(got - ref)
        # a pedestrian sitting within float rounding of a cell boundary can land in the neighbouring cell in a
        # free-running rollout; such tracks are COUNTED, everything else must agree to 1e-4
        per_track = np.nanmax(np.where(np.isnan(d), 0.0, d), axis=(0, 2))
        flipped = int((per_track > 1e-4).sum())
        print("%s%s: %d of %d tracks beyond 1e-4 (max %.2e), median %.2e" % (kind, key, flipped, M, float(np.nanmax(d)), float(np.nanmedian(d))))
        assert np.nanmedian(d) < 1e-5
        assert flipped == 0, (kind, key, flipped, float(np.nanmax(d)))
