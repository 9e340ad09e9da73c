"""Non-grid interaction modules HiddenStateMLPPooling, NearestNeighborMLP and AttentionMLPPooling (SURVEY.md 8f rank 4;
reference lstm/non_gridbased_pooling.py:150-239, :64-147 and :242-351) against vectors the unmodified reference produced
(oracle/make_nongrid_golden.py): the numpy oracle on CPU, the CUDA kernel behind the plug and inside
LSTM.forward on the GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import lstm_oracle as O
from oracle.make_nongrid_golden import ATTN_KINDS, KINDS, NN_KINDS, NN_LSTM_KINDS, TRAJ_KINDS, plug_inputs, scene_inputs

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "nongrid_golden.npz"))


ALL_KINDS = KINDS + NN_KINDS + ATTN_KINDS + NN_LSTM_KINDS + TRAJ_KINDS
STATEFUL_KINDS = NN_LSTM_KINDS + TRAJ_KINDS


def _pool(kind):
    from trajnetplusplusbaselines_b200.lstm import (AttentionMLPPooling, HiddenStateMLPPooling, NearestNeighborLSTM, NearestNeighborMLP,
                                                    TrajectronPooling)
    if kind in O.TRAJ_SPECS:
        return TrajectronPooling(**O.TRAJ_SPECS[kind])
    if kind in O.NN_LSTM_SPECS:
        return NearestNeighborLSTM(**O.NN_LSTM_SPECS[kind])
    if kind in O.NN_SPECS:
        return NearestNeighborMLP(**O.NN_SPECS[kind])
    if kind in O.ATTN_SPECS:
        return AttentionMLPPooling(**O.ATTN_SPECS[kind])
    return HiddenStateMLPPooling(**O.NONGRID_SPECS[kind])


@pytest.mark.parametrize("kind", ALL_KINDS)
def test_oracle_matches_reference_vectors(kind):
    W = O.random_weights(kind, seed=13)
    cfg = O.pool_config(kind)
    hid, obs1, obs2 = plug_inputs()
    if kind in STATEFUL_KINDS:      # stateful plug: two calls after a reset
        n = obs2.shape[0] * obs2.shape[1]
        st = {"h": np.zeros((n, cfg.hidden_dim), np.float32), "c": np.zeros((n, cfg.hidden_dim), np.float32)}
        assert np.abs(O.pool_forward(cfg, W, hid, obs1, obs2, state=st) - GOLD[kind + "/plug"]).max() < 1e-5
        assert np.abs(O.pool_forward(cfg, W, hid, obs2, obs2 + (obs2 - obs1), state=st) - GOLD[kind + "/plug2"]).max() < 1e-5
    else:
        assert np.abs(O.pool_forward(cfg, W, hid, obs1, obs2) - GOLD[kind + "/plug"]).max() < 1e-5
    xy, bs = scene_inputs()
    rel, pred = O.forward(W, cfg, xy[:9], bs, n_predict=12)
    _, pred_t = O.forward(W, cfg, xy[:9], bs, prediction_truth=xy[9:20])
    for got, key in ((rel, "/rel_free"), (pred, "/pred_free"), (pred_t, "/pred_teacher")):
        ref = GOLD[kind + key]
        assert (np.isnan(got) == np.isnan(ref)).all()
        assert np.nanmax(np.abs(got - ref)) < 2e-5


def test_state_dict_keys_match_reference_layout():
    from trajnetplusplusbaselines_b200.lstm import LSTM
    for kind in ALL_KINDS:
        model = LSTM(pool=_pool(kind))
        W = O.random_weights(kind, seed=13)           # keys / shapes checked against the reference by the generator
        sd = model.state_dict()
        assert set(sd.keys()) == set(W.keys())
        for k, v in W.items():
            assert tuple(sd[k].shape) == v.shape, k


def _model(kind):
    from trajnetplusplusbaselines_b200.lstm import LSTM
    model = LSTM(pool=_pool(kind))
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in O.random_weights(kind, seed=13).items()}, strict=True)
    return model.cuda().eval()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ALL_KINDS)
def test_cuda_plug_matches_reference_vectors(kind):
    from trajnetplusplusbaselines_b200 import _lib
    model = _model(kind)
    hid, obs1, obs2 = plug_inputs()
    before = _lib.load().tb2_launch_count()
    out = model.pool(torch.from_numpy(hid).cuda(), torch.from_numpy(obs1).cuda(), torch.from_numpy(obs2).cuda())
    assert _lib.load().tb2_launch_count() > before
    ref = GOLD[kind + "/plug"]
    assert out.shape == ref.shape
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-4 * max(1.0, float(np.abs(ref).max()))      # fp32 order of the 128-term sums
    if kind in STATEFUL_KINDS:      # the state advanced: second call, then the same pair again after a reset
        o2 = torch.from_numpy(obs2).cuda()
        o3 = torch.from_numpy(obs2 + (obs2 - obs1)).cuda()
        out2 = model.pool(torch.from_numpy(hid).cuda(), o2, o3)
        ref2 = GOLD[kind + "/plug2"]
        assert np.abs(out2.cpu().numpy() - ref2).max() < 1e-4 * max(1.0, float(np.abs(ref2).max()))
        model.pool.reset(obs2.shape[0] * obs2.shape[1], obs2.shape[1] - 1, device=o2.device)
        again = model.pool(torch.from_numpy(hid).cuda(), torch.from_numpy(obs1).cuda(), o2)
        assert np.abs(again.cpu().numpy() - ref).max() < 1e-4 * max(1.0, float(np.abs(ref).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ALL_KINDS)
def test_cuda_forward_matches_reference_vectors(kind):
    model = _model(kind)
    xy, bs = scene_inputs()
    M = xy.shape[1]
    with torch.no_grad():
        rel, pred = model(torch.from_numpy(xy[:9]), torch.zeros(M, 2), torch.from_numpy(bs), n_predict=12)
        _, pred_t = model(torch.from_numpy(xy[:9]), torch.zeros(M, 2), torch.from_numpy(bs),
                          prediction_truth=torch.from_numpy(xy[9:20]).clone())
    for got, key in ((rel, "/rel_free"), (pred, "/pred_free"), (pred_t, "/pred_teacher")):
        ref = GOLD[kind + key]
        got = got.numpy()
        assert (np.isnan(got) == np.isnan(ref)).all()
        assert np.nanmax(np.abs(got - ref)) < 1e-4, (kind, key, float(np.nanmax(np.abs(got - ref))))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hiddenstatemlp", "nn", "attentionmlp", "nn_lstm", "traj_pool"])
def test_cuda_baseline_shape_vs_oracle_and_training_raises(kind):
    """256-d pooling at N = 20, T = 9 + 12 on 48 scenes vs the oracle; training is inference-only."""
    model = _model(kind)
    xy, bs = O.synthetic_scenes(48, 20, seed=3, nan_tracks=True)
    M = xy.shape[1]
    with torch.no_grad():
        _, pred = model(torch.from_numpy(xy[:9]), torch.zeros(M, 2), torch.from_numpy(bs), n_predict=12)
    _, pred_o = O.forward(O.random_weights(kind, seed=13), O.pool_config(kind), xy[:9], bs, n_predict=12)
    pred = pred.numpy()
    assert (np.isnan(pred) == np.isnan(pred_o)).all()
    assert np.nanmax(np.abs(pred - pred_o)) < 1e-4
    model.train()
    with pytest.raises(NotImplementedError):
        rel, _ = model(torch.from_numpy(xy[:9]).cuda(), torch.zeros(M, 2), torch.from_numpy(bs),
                       prediction_truth=torch.from_numpy(xy[9:20]).cuda())
        rel.sum().backward()
