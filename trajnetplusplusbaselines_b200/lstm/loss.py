"""Losses of trajnetbaselines/lstm/loss.py (PredictionLoss :6-91, L2Loss :93-135).

Placeholder until the CUDA training path lands; see lstm/training.py.
"""
import torch


class PredictionLoss(torch.nn.Module):
    def __init__(self, keep_batch_dim=False, background_rate=0.2, col_wt=0.0, col_distance=0.2):
        super().__init__()
        self.keep_batch_dim = keep_batch_dim
        self.background_rate = background_rate
        self.loss_multiplier = 1
        self.col_wt = col_wt
        self.col_distance = col_distance

    def forward(self, inputs, targets, batch_split, positions=None):
        raise NotImplementedError("PredictionLoss CUDA kernel not built yet")


class L2Loss(torch.nn.Module):
    def __init__(self, keep_batch_dim=False, col_wt=0.0, col_distance=0.2):
        super().__init__()
        self.keep_batch_dim = keep_batch_dim
        self.loss_multiplier = 100
        self.col_wt = col_wt
        self.col_distance = col_distance

    def forward(self, inputs, targets, batch_split, positions=None):
        raise NotImplementedError("L2Loss CUDA kernel not built yet")
