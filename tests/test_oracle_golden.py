"""The numpy oracle against fixtures generated from the reference (oracle/make_golden.py) and
against the reference's own (adapted) golden vectors, SURVEY.md section 4.  CPU only."""
import numpy as np
import pytest

from oracle import lstm_oracle as O
from oracle.make_golden import CASES


def _close(a, b, tol):
    assert a.shape == b.shape
    assert (np.isnan(a) == np.isnan(b)).all()
    return np.nanmax(np.abs(a - b)) <= tol if a.size else True


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_forward_matches_reference_fixture(golden, case):
    name, kind, B, N, ragged, nan_tracks, dseed, wseed, wscale = case
    xy, bs = O.synthetic_scenes(B, N, seed=dseed, ragged=ragged, nan_tracks=nan_tracks)
    W = O.random_weights(kind, seed=wseed, scale=wscale)
    cfg = O.pool_config(kind)
    rel_f, pred_f = O.forward(W, cfg, xy[:9], bs, n_predict=12)
    rel_t, pred_t = O.forward(W, cfg, xy[:9], bs, prediction_truth=xy[9:20])
    # fp32 on both sides, different BLAS summation order: 2e-5 m is ~40x headroom on what we see
    assert _close(rel_f, golden[name + "/rel_free"], 2e-5)
    assert _close(pred_f, golden[name + "/pred_free"], 2e-5)
    assert _close(rel_t, golden[name + "/rel_teacher"], 2e-5)
    assert _close(pred_t, golden[name + "/pred_teacher"], 2e-5)


@pytest.mark.parametrize("kind", ["social", "directional", "occupancy_front"])
def test_grid_cells_bit_exact(golden, kind):
    obs = golden["cells_%s/obs" % kind]
    cells, inr = O.grid_cells(obs, O.pool_config(kind))
    assert (inr == golden["cells_%s/in_range" % kind]).all()
    assert (cells.astype(np.int32) == golden["cells_%s/cells" % kind]).all()


def _grid(obs1, obs2, **kw):
    cfg = O.PoolConfig(embedding_arch="None", **kw)
    o1 = np.array([obs1], dtype=np.float32)
    o2 = np.array([obs2], dtype=np.float32)
    h = np.zeros((1, len(obs1), 128), dtype=np.float32)
    return O.pool_forward(cfg, {}, h, o1, o2)


def test_reference_golden_simple_grid(golden):
    # reference tests/test_pooling.py:9-22 (adapted to the HEAD API)
    g = _grid([[0, 0], [-1, -1]], [[0, 0], [-1, -1]], n=2, pool_size=4, blur_size=3, cell_side=2.0)
    assert np.allclose(g, [[1, 0, 0, 0], [0, 0, 0, 1]], atol=1e-6)
    assert np.allclose(g, golden["sec4/simple_grid"], atol=1e-6)


def test_reference_golden_midpoint(golden):
    # reference tests/test_pooling.py:65-83
    g = _grid([[0, 0], [-1, 0]], [[0, 0], [-1, 0]], n=2, pool_size=100, blur_size=99, cell_side=2.0)
    assert np.allclose(g, [[0.5, 0.5, 0, 0], [0, 0, 0.5, 0.5]], atol=0.01)
    # 99x99 blur + 100x100 window sum in fp32: summation order differs from torch's pooling
    assert np.allclose(g, golden["sec4/simple_grid_midpoint"], atol=1e-4)


def test_reference_golden_nan(golden):
    # reference tests/test_pooling.py:86-99
    nan = float("nan")
    g = _grid([[0, 0], [nan, nan]], [[0, 0], [nan, nan]], n=2, cell_side=2.0)
    assert (g[0] == 0).all()
    assert np.array_equal(g, golden["sec4/nan"])


def test_reference_golden_directional(golden):
    # reference tests/test_pooling.py:45-62; HEAD stores the RELATIVE velocity (-/+0.2, not 0.1)
    g = _grid([[0, 0], [-1, -1]], [[0.1, 0.1], [-1.1, -1.1]], n=2, pool_size=4, cell_side=2.0,
              type_="directional")
    assert np.allclose(g, golden["sec4/directional"], atol=1e-6)
    assert np.isclose(np.abs(g).max(), 0.2, atol=1e-6)


def test_reference_golden_loss(golden):
    # reference tests/test_lstm_loss.py:12-25: -log(0.01 + 0.2 N(0;0,3) + 0.79 N(0;0,1))
    p = np.array([[[0.0, 0.0, 1.0, 1.0, 0.0]]], dtype=np.float32)
    t = np.array([[[0.0, 0.0]]], dtype=np.float32)
    v = O.prediction_loss(p, t, [0, 1])
    assert abs(float(v) - float(golden["sec4/loss_simple"][0])) < 1e-6


def test_cell_zero_clobber_semantics():
    """Out-of-range neighbours overwrite cell 0 with `constant`, in ascending-j order
    (gridbased_pooling.py:281-293)."""
    cfg = O.PoolConfig(type_="occupancy", n=4, cell_side=1.0, embedding_arch="None")
    # ped 0 at origin; ped 1 lands in cell 0 (ox, oy in [0,1)); ped 2 far away (out of range)
    obs = np.array([[[0, 0], [-1.5, -1.5], [50, 50]]], dtype=np.float32)
    g = O.occupancy_grid(obs, None, cfg).reshape(3, -1)
    assert g[0, 0] == 0          # in-range writer (j=1) clobbered by the later out-of-range j=2
    obs2 = obs[:, [0, 2, 1]]     # swap: out-of-range first, in-range last -> survives
    g2 = O.occupancy_grid(obs2, None, cfg).reshape(3, -1)
    assert g2[0, 0] == 1


def test_padding_clobbers_cell_zero():
    """A scene smaller than the batch maximum is NaN-padded (lstm.py:31-40) and the padded slots
    are trailing out-of-range writers of cell 0."""
    cfg = O.PoolConfig(type_="occupancy", n=4, cell_side=1.0, embedding_arch="None")
    nan = np.nan
    obs = np.array([[[0, 0], [-1.5, -1.5], [nan, nan]]], dtype=np.float32)
    g = O.occupancy_grid(obs, None, cfg).reshape(3, -1)
    assert g[0, 0] == 0
    g_np = O.occupancy_grid(obs[:, :2], None, cfg).reshape(2, -1)
    assert g_np[0, 0] == 1
