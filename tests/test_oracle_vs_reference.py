"""Oracle vs the reference imported live from /root/reference (build container only)."""
import numpy as np
import pytest

from oracle import lstm_oracle as O

pytestmark = pytest.mark.needs_reference


@pytest.mark.parametrize("kind", ["vanilla", "directional", "social_small", "occupancy_front", "directional_const"])
@pytest.mark.parametrize("variant", ["plain", "ragged_nan"])
def test_forward_live(kind, variant):
    import torch
    from oracle.ref_shim import import_reference
    import_reference()
    from oracle.make_golden import build_reference_model
    ragged = variant == "ragged_nan"
    xy, bs = O.synthetic_scenes(7, 9, seed=123, ragged=ragged, nan_tracks=ragged)
    W = O.random_weights(kind, seed=5)
    model = build_reference_model(kind, W)
    M = xy.shape[1]
    with torch.no_grad():
        rel, pred = model(torch.from_numpy(xy[:9]), torch.zeros(M, 2), torch.from_numpy(bs), n_predict=12)
    rel_o, pred_o = O.forward(W, O.pool_config(kind), xy[:9], bs, n_predict=12)
    assert (np.isnan(rel.numpy()) == np.isnan(rel_o)).all()
    assert np.nanmax(np.abs(rel.numpy() - rel_o)) < 2e-5
    assert np.nanmax(np.abs(pred.numpy() - pred_o)) < 2e-5


def test_social_full_config_live():
    import torch
    from oracle.ref_shim import import_reference
    import_reference()
    from oracle.make_golden import build_reference_model
    xy, bs = O.synthetic_scenes(4, 8, seed=7)
    W = O.random_weights("social", seed=3)
    model = build_reference_model("social", W)
    M = xy.shape[1]
    with torch.no_grad():
        rel, pred = model(torch.from_numpy(xy[:9]), torch.zeros(M, 2), torch.from_numpy(bs), n_predict=12)
    rel_o, pred_o = O.forward(W, O.pool_config("social"), xy[:9], bs, n_predict=12)
    assert np.nanmax(np.abs(pred.numpy() - pred_o)) < 2e-5
