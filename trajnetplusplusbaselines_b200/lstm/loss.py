"""Losses with the reference's API (trajnetbaselines/lstm/loss.py: PredictionLoss :6-91, L2Loss
:93-135).

The loss acts on [pred_length x batch_size] primary rows only (a few thousand elements); it is
expressed with torch tensor ops on whatever device the predictions live on, so autograd chains it
into the CUDA BPTT of lstm/training.py.  (SURVEY.md build plan step 6: "loss kept in torch first".)
"""
import math

import torch


class PredictionLoss(torch.nn.Module):
    """2D Gaussian with a flat background: -log(0.01 + bg N(x|mu, 3) + (0.99 - bg) N(x|mu, sigma, rho))."""

    def __init__(self, keep_batch_dim=False, background_rate=0.2, col_wt=0.0, col_distance=0.2):
        super().__init__()
        self.keep_batch_dim = keep_batch_dim
        self.background_rate = background_rate
        self.loss_multiplier = 1
        self.col_wt = col_wt
        self.col_distance = col_distance
        if self.col_wt:
            raise NotImplementedError("auxiliary collision loss (col_wt != 0) is not built")

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
        return torch.mean(values) * self.loss_multiplier


class L2Loss(torch.nn.Module):
    """Deterministic variant (loss.py:93-135): 100 x MSE on the primaries' mean prediction."""

    def __init__(self, keep_batch_dim=False, col_wt=0.0, col_distance=0.2):
        super().__init__()
        self.keep_batch_dim = keep_batch_dim
        self.loss_multiplier = 100
        self.col_wt = col_wt
        self.col_distance = col_distance
        if self.col_wt:
            raise NotImplementedError("auxiliary collision loss (col_wt != 0) is not built")

    def forward(self, inputs, targets, batch_split, positions=None):
        batch_split = torch.as_tensor(batch_split)
        prim = batch_split[:-1].to(inputs.device)
        targets = targets.to(inputs.device)[:, prim]
        inputs = inputs[:, prim]
        loss = (inputs[:, :, :2] - targets) ** 2
        if self.keep_batch_dim:
            return loss.mean(dim=0).mean(dim=1) * self.loss_multiplier
        return torch.mean(loss) * self.loss_multiplier
