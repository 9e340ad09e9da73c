"""Training losses vs vectors produced by the unmodified reference (oracle/make_loss_golden.py):
PredictionLoss / L2Loss with and without the collision term (lstm/loss.py:52-162).
CPU: the torch-op expression; GPU: the fused kernels of csrc/loss.cu through the C ABI."""
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_golden.npz"))
CASES = ["uniform", "ragged", "big"]


def _run(case, col_wt, which, device):
    from trajnetplusplusbaselines_b200.lstm import L2Loss, PredictionLoss
    inputs = torch.from_numpy(GOLD[case + "/inputs"].copy()).to(device).requires_grad_(True)
    pos = torch.from_numpy(GOLD[case + "/pos"].copy()).to(device).requires_grad_(True)
    targets = torch.from_numpy(GOLD[case + "/targets"]).to(device)
    bs = torch.from_numpy(GOLD[case + "/bs"])
    crit = (PredictionLoss if which == "pl" else L2Loss)(col_wt=col_wt, col_distance=0.2)
    loss = crit(inputs, targets, bs, (pos * 1.0) if col_wt else None)
    loss.backward()
    dpos = pos.grad.cpu().numpy() if pos.grad is not None else np.zeros(pos.shape, np.float32)
    return float(loss.item()), inputs.grad.cpu().numpy(), dpos


def _check(case, col_wt, which, device):
    loss, dinputs, dpos = _run(case, col_wt, which, device)
    key = "%s/col%d/%s" % (case, int(col_wt), which)
    ref = float(GOLD[key + "/loss"][0])
    assert abs(loss - ref) <= 2e-5 * max(1.0, abs(ref)), (loss, ref)
    gi, gp = GOLD[key + "/dinputs"], GOLD[key + "/dpos"]
    assert np.abs(dinputs - gi).max() <= 2e-5 * max(1.0, np.abs(gi).max())
    assert np.abs(dpos - gp).max() <= 2e-5 * max(1.0, np.abs(gp).max())
    if col_wt:
        assert np.abs(gp).max() > 0          # the fixture really has collisions


@pytest.mark.parametrize("which", ["pl", "l2"])
@pytest.mark.parametrize("col_wt", [0.0, 10.0])
@pytest.mark.parametrize("case", CASES)
def test_losses_cpu_match_reference(case, col_wt, which):
    _check(case, col_wt, which, "cpu")


@pytest.mark.parametrize("case", CASES)
def test_keep_batch_dim_cpu(case):
    from trajnetplusplusbaselines_b200.lstm import PredictionLoss
    out = PredictionLoss(keep_batch_dim=True)(torch.from_numpy(GOLD[case + "/inputs"]),
                                              torch.from_numpy(GOLD[case + "/targets"]),
                                              torch.from_numpy(GOLD[case + "/bs"]))
    assert np.abs(out.numpy() - GOLD[case + "/col0/pl/keep_batch"]).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["pl", "l2"])
@pytest.mark.parametrize("col_wt", [0.0, 10.0])
@pytest.mark.parametrize("case", CASES)
def test_losses_cuda_match_reference(case, col_wt, which):
    from trajnetplusplusbaselines_b200 import _lib
    lib = _lib.load()
    before = lib.tb2_launch_count()
    _check(case, col_wt, which, "cuda")
    if which == "pl" or col_wt:
        assert lib.tb2_launch_count() > before      # the fused kernels ran, not a torch fallback


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_keep_batch_dim_cuda(case):
    from trajnetplusplusbaselines_b200.lstm import PredictionLoss
    out = PredictionLoss(keep_batch_dim=True)(torch.from_numpy(GOLD[case + "/inputs"]).cuda(),
                                              torch.from_numpy(GOLD[case + "/targets"]).cuda(),
                                              torch.from_numpy(GOLD[case + "/bs"]))
    assert np.abs(out.cpu().numpy() - GOLD[case + "/col0/pl/keep_batch"]).max() < 1e-5
