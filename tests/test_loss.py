"""Training losses vs vectors produced by the unmodified reference (oracle/make_loss_golden.py):
PredictionLoss / L2Loss with and without the collision term (lstm/loss.py:52-162).
CPU: the torch restatement kept under tests/ (torch_ref.py) is pinned to the reference's values and
gradients, and the product refuses CPU tensors; GPU: the fused kernels of csrc/loss.cu through the C ABI."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import torch_ref as TR  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_golden.npz"))
CASES = ["uniform", "ragged", "big"]


def _tensors(case, device):
    inputs = torch.from_numpy(GOLD[case + "/inputs"].copy()).to(device).requires_grad_(True)
    pos = torch.from_numpy(GOLD[case + "/pos"].copy()).to(device).requires_grad_(True)
    targets = torch.from_numpy(GOLD[case + "/targets"]).to(device)
    bs = torch.from_numpy(GOLD[case + "/bs"])
    return inputs, pos, targets, bs


def _compare(case, col_wt, which, loss, inputs, pos):
    loss.backward()
    dinputs = inputs.grad.cpu().numpy()
    dpos = pos.grad.cpu().numpy() if pos.grad is not None else np.zeros(pos.shape, np.float32)
    key = "%s/col%d/%s" % (case, int(col_wt), which)
    ref = float(GOLD[key + "/loss"][0])
    assert abs(float(loss.item()) - ref) <= 2e-5 * max(1.0, abs(ref)), (float(loss.item()), ref)
    gi, gp = GOLD[key + "/dinputs"], GOLD[key + "/dpos"]
    assert np.abs(dinputs - gi).max() <= 2e-5 * max(1.0, np.abs(gi).max())
    assert np.abs(dpos - gp).max() <= 2e-5 * max(1.0, np.abs(gp).max())
    if col_wt:
        assert np.abs(gp).max() > 0          # the fixture really has collisions


@pytest.mark.parametrize("which", ["pl", "l2"])
@pytest.mark.parametrize("col_wt", [0.0, 10.0])
@pytest.mark.parametrize("case", CASES)
def test_torch_restatement_matches_reference(case, col_wt, which):
    """tests/torch_ref.py (the checker of the CUDA backward) against the reference's goldens."""
    inputs, pos, targets, bs = _tensors(case, "cpu")
    mult = 1 if which == "pl" else 100
    loss = TR.prediction_loss(inputs, targets, bs) if which == "pl" else TR.l2_loss(inputs, targets, bs)
    if col_wt:
        loss = loss + TR.collision_loss(pos * 1.0, bs, col_wt, 0.2) * mult
    _compare(case, col_wt, which, loss, inputs, pos)


@pytest.mark.parametrize("case", CASES)
def test_torch_restatement_keep_batch_dim(case):
    v = TR.prediction_loss_values(torch.from_numpy(GOLD[case + "/inputs"]), torch.from_numpy(GOLD[case + "/targets"]),
                                  torch.from_numpy(GOLD[case + "/bs"]))
    assert np.abs(v.mean(dim=0).numpy() - GOLD[case + "/col0/pl/keep_batch"]).max() < 1e-5


def test_losses_refuse_cpu_tensors():
    """No CPU / torch fallback in the product: host tensors raise (whatever the reason: no library,
    no device, or the explicit device check)."""
    from trajnetplusplusbaselines_b200.lstm import L2Loss, PredictionLoss
    from trajnetplusplusbaselines_b200.lstm.loss import collision_loss
    inputs, pos, targets, bs = _tensors("uniform", "cpu")
    for crit in (PredictionLoss(), L2Loss()):
        with pytest.raises(Exception):
            crit(inputs, targets, bs)
    with pytest.raises(Exception):
        collision_loss(pos, bs.tolist())


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["pl", "l2"])
@pytest.mark.parametrize("col_wt", [0.0, 10.0])
@pytest.mark.parametrize("case", CASES)
def test_losses_cuda_match_reference(case, col_wt, which):
    from trajnetplusplusbaselines_b200 import _lib
    from trajnetplusplusbaselines_b200.lstm import L2Loss, PredictionLoss
    lib = _lib.load()
    before = lib.tb2_launch_count()
    inputs, pos, targets, bs = _tensors(case, "cuda")
    crit = (PredictionLoss if which == "pl" else L2Loss)(col_wt=col_wt, col_distance=0.2)
    loss = crit(inputs, targets, bs, (pos * 1.0) if col_wt else None)
    _compare(case, col_wt, which, loss, inputs, pos)
    assert lib.tb2_launch_count() >= before + (2 if col_wt else 1)      # the fused kernels ran


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_keep_batch_dim_cuda(case):
    from trajnetplusplusbaselines_b200.lstm import PredictionLoss
    out = PredictionLoss(keep_batch_dim=True)(torch.from_numpy(GOLD[case + "/inputs"]).cuda(),
                                              torch.from_numpy(GOLD[case + "/targets"]).cuda(),
                                              torch.from_numpy(GOLD[case + "/bs"]))
    assert np.abs(out.cpu().numpy() - GOLD[case + "/col0/pl/keep_batch"]).max() < 1e-5
