"""Losses with the reference's API (trajnetbaselines/lstm/loss.py: PredictionLoss :6-91, L2Loss
:93-135).

The loss acts on [pred_length x batch_size] primary rows only.  For CUDA fp32 predictions the
per-(frame, scene) value and its analytic derivative come from one kernel each
(tb2_prediction_loss / tb2_collision_loss, csrc/loss.cu) behind a torch.autograd.Function, so
autograd chains them into the CUDA BPTT of lstm/training.py; tensors on the CPU go through the
same expression written with torch ops (host-side checks of the kernels, no device work).
"""
import ctypes
import math

import torch

from .. import _lib
from ..engine import LayoutCache, _ptr, _stream

_layouts = LayoutCache(capacity=8)


class _PredictionLossFn(torch.autograd.Function):
    """values [T, B] of the primaries; backward scatters d value / d inputs into [T, M, 5]."""

    @staticmethod
    def forward(ctx, inputs, targets, prim, background_rate):
        lib = _lib.load()
        inputs = inputs.contiguous()
        targets = targets.contiguous()
        T, M = int(inputs.shape[0]), int(inputs.shape[1])
        B = int(prim.numel())
        values = torch.empty((T, B), dtype=torch.float32, device=inputs.device)
        dinputs = torch.empty((T, B, 5), dtype=torch.float32, device=inputs.device)
        with torch.cuda.device(inputs.device):
            _lib.check(lib.tb2_prediction_loss(_ptr(inputs), _ptr(targets), _ptr(prim), T, M, B,
                                               ctypes.c_float(background_rate), _ptr(values), _ptr(dinputs),
                                               _stream(inputs.device)))
        ctx.save_for_backward(dinputs, prim)
        ctx.shape = (T, M)
        return values

    @staticmethod
    def backward(ctx, grad_values):
        dinputs, prim = ctx.saved_tensors
        T, M = ctx.shape
        grad = torch.zeros((T, M, 5), dtype=torch.float32, device=dinputs.device)
        grad[:, prim.long()] = dinputs * grad_values.unsqueeze(-1)
        return grad, None, None, None


class _CollisionLossFn(torch.autograd.Function):
    """CollisionLoss (loss.py:138-162) summed over frames, scenes and colliding neighbours."""

    @staticmethod
    def forward(ctx, positions, batch_split, col_wt, col_distance):
        lib = _lib.load()
        positions = positions.contiguous()
        layout = _layouts.get(batch_split)
        T, M = int(positions.shape[0]), int(positions.shape[1])
        B = layout.num_scenes
        per = torch.empty((T, B), dtype=torch.float32, device=positions.device)
        dprim = torch.empty((T, B, 2), dtype=torch.float32, device=positions.device)
        with torch.cuda.device(positions.device):
            _lib.check(lib.tb2_collision_loss(layout.handle, _ptr(positions), T, ctypes.c_float(col_wt),
                                              ctypes.c_float(col_distance), _ptr(per), _ptr(dprim),
                                              _stream(positions.device)))
        ctx.save_for_backward(dprim)
        ctx.prim = [int(v) for v in batch_split[:-1]]
        ctx.shape = (T, M)
        return per.sum()

    @staticmethod
    def backward(ctx, grad_out):
        (dprim,) = ctx.saved_tensors
        T, M = ctx.shape
        grad = torch.zeros((T, M, 2), dtype=torch.float32, device=dprim.device)
        grad[:, torch.as_tensor(ctx.prim, device=dprim.device)] = dprim * grad_out
        return grad, None, None, None


def collision_loss(predictions, batch_split, col_wt=10.0, col_distance=0.2):
    """loss.py:138-162.  predictions [pred_length, num_tracks, 2]: the primary of each scene is
    penalised for coming within col_distance of a neighbour (neighbours are constants; NaN
    coordinates count as -1000)."""
    batch_split = [int(v) for v in batch_split]
    if predictions.is_cuda and predictions.dtype == torch.float32:
        return _CollisionLossFn.apply(predictions[..., :2], batch_split, float(col_wt), float(col_distance))
    pos = torch.where(torch.isnan(predictions[..., :2]), torch.full_like(predictions[..., :2], -1000.0),
                      predictions[..., :2])
    sizes = torch.as_tensor([b - a for a, b in zip(batch_split[:-1], batch_split[1:])])
    prim_of_row = torch.repeat_interleave(torch.as_tensor(batch_split[:-1]), sizes)
    is_neigh = torch.ones(batch_split[-1], dtype=torch.bool)
    is_neigh[torch.as_tensor(batch_split[:-1])] = False
    dist = torch.norm(pos[:, prim_of_row] - pos.detach(), dim=-1)[:, is_neigh]
    hit = (dist <= col_distance).detach()
    return col_wt * (1 - dist[hit] / col_distance).sum()


class PredictionLoss(torch.nn.Module):
    """2D Gaussian with a flat background: -log(0.01 + bg N(x|mu, 3) + (0.99 - bg) N(x|mu, sigma, rho))."""

    def __init__(self, keep_batch_dim=False, background_rate=0.2, col_wt=0.0, col_distance=0.2):
        super().__init__()
        self.keep_batch_dim = keep_batch_dim
        self.background_rate = background_rate
        self.loss_multiplier = 1
        self.col_wt = col_wt
        self.col_distance = col_distance

    @staticmethod
    def gaussian_2d(mu1mu2s1s2rho, x1x2):
        """loss.py:24-50."""
        x1, x2 = x1x2[:, 0], x1x2[:, 1]
        mu1, mu2, s1, s2, rho = (mu1mu2s1s2rho[:, i] for i in range(5))
        norm1 = x1 - mu1
        norm2 = x2 - mu2
        sigma1sigma2 = s1 * s2
        z = (norm1 / s1) ** 2 + (norm2 / s2) ** 2 - 2 * rho * norm1 * norm2 / sigma1sigma2
        numerator = torch.exp(-z / (2 * (1 - rho ** 2)))
        denominator = 2 * math.pi * sigma1sigma2 * torch.sqrt(1 - rho ** 2)
        return numerator / denominator

    def forward(self, inputs, targets, batch_split, positions=None):
        """inputs [pred_length, num_tracks, 5], targets [pred_length, num_tracks, 2] (loss.py:52-91)."""
        batch_split = torch.as_tensor(batch_split)
        pred_length, batch_size = targets.size(0), batch_split[:-1].size(0)
        col_loss = 0
        if self.col_wt:
            assert positions is not None, "Prediction positions required to calculate collision loss"
            col_loss = collision_loss(positions, batch_split.tolist(), self.col_wt, self.col_distance)
        if inputs.is_cuda and inputs.dtype == torch.float32:
            prim = batch_split[:-1].to(device=inputs.device, dtype=torch.int32)
            values = _PredictionLossFn.apply(inputs, targets.to(device=inputs.device, dtype=torch.float32), prim,
                                             float(self.background_rate))
            if self.keep_batch_dim:
                return values.mean(dim=0) * self.loss_multiplier
            if self.col_wt:
                return torch.mean(values) * self.loss_multiplier + col_loss * self.loss_multiplier
            return torch.mean(values) * self.loss_multiplier
        prim = batch_split[:-1].to(inputs.device)
        targets = targets.to(inputs.device)[:, prim].reshape(-1, 2)
        inputs = inputs[:, prim].reshape(-1, 5)
        inputs_bg = inputs.clone()
        inputs_bg[:, 2] = 3.0
        inputs_bg[:, 3] = 3.0
        inputs_bg[:, 4] = 0.0
        values = -torch.log(
            0.01 +
            self.background_rate * self.gaussian_2d(inputs_bg, targets) +
            (0.99 - self.background_rate) * self.gaussian_2d(inputs, targets))
        if self.keep_batch_dim:
            return values.reshape(pred_length, batch_size).mean(dim=0) * self.loss_multiplier
        if self.col_wt:
            return torch.mean(values) * self.loss_multiplier + col_loss * self.loss_multiplier
        return torch.mean(values) * self.loss_multiplier


class L2Loss(torch.nn.Module):
    """Deterministic variant (loss.py:93-135): 100 x MSE on the primaries' mean prediction."""

    def __init__(self, keep_batch_dim=False, col_wt=0.0, col_distance=0.2):
        super().__init__()
        self.keep_batch_dim = keep_batch_dim
        self.loss_multiplier = 100
        self.col_wt = col_wt
        self.col_distance = col_distance

    def forward(self, inputs, targets, batch_split, positions=None):
        batch_split = torch.as_tensor(batch_split)
        col_loss = 0
        if self.col_wt:
            assert positions is not None, "Prediction positions required to calculate collision loss"
            col_loss = collision_loss(positions, batch_split.tolist(), self.col_wt, self.col_distance)
        prim = batch_split[:-1].to(inputs.device)
        targets = targets.to(inputs.device)[:, prim]
        inputs = inputs[:, prim]
        loss = (inputs[:, :, :2] - targets) ** 2
        if self.keep_batch_dim:
            return loss.mean(dim=0).mean(dim=1) * self.loss_multiplier
        if self.col_wt:
            return torch.mean(loss) * self.loss_multiplier + col_loss * self.loss_multiplier
        return torch.mean(loss) * self.loss_multiplier
