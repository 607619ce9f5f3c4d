"""Exp-normal radial basis with cosine cutoff (MD17 models; drop-in for ``nets/expnorm_rbf.py``)."""
from __future__ import annotations

import math

import torch
from torch import nn


class CosineCutoff(nn.Module):
    def __init__(self, cutoff_lower: float = 0.0, cutoff_upper: float = 5.0):
        super().__init__()
        self.cutoff_lower = cutoff_lower
        self.cutoff_upper = cutoff_upper

    def forward(self, distances):
        lo, hi = self.cutoff_lower, self.cutoff_upper
        if lo > 0:
            c = 0.5 * (torch.cos(math.pi * (2 * (distances - lo) / (hi - lo) + 1.0)) + 1.0)
            return c * (distances < hi).float() * (distances > lo).float()
        c = 0.5 * (torch.cos(distances * math.pi / hi) + 1.0)
        return c * (distances < hi).float()


class ExpNormalSmearing(nn.Module):
    def __init__(self, cutoff_lower: float = 0.0, cutoff_upper: float = 5.0, num_rbf: int = 50, trainable: bool = False):
        super().__init__()
        self.cutoff_lower, self.cutoff_upper = cutoff_lower, cutoff_upper
        self.num_rbf, self.trainable = num_rbf, trainable
        self.cutoff_fn = CosineCutoff(0, cutoff_upper)
        self.alpha = 5.0 / (cutoff_upper - cutoff_lower)
        means, betas = self._initial_params()
        if trainable:
            self.register_parameter("means", nn.Parameter(means))
            self.register_parameter("betas", nn.Parameter(betas))
        else:
            self.register_buffer("means", means)
            self.register_buffer("betas", betas)

    def _initial_params(self):
        start = torch.exp(torch.scalar_tensor(-self.cutoff_upper + self.cutoff_lower))
        means = torch.linspace(start, 1, self.num_rbf)
        betas = torch.tensor([(2 / self.num_rbf * (1 - start)) ** -2] * self.num_rbf)
        return means, betas

    def reset_parameters(self):
        means, betas = self._initial_params()
        self.means.data.copy_(means)
        self.betas.data.copy_(betas)

    def forward(self, dist):
        if self.cutoff_lower == 0 and dist.is_cuda and dist.dtype == torch.float32 and dist.dim() == 1:
            from .. import ops          # one kernel each way (ops.ExpNormalRbf); second order via the torch statement
            return ops.expnorm_rbf(dist, self.means, self.betas, self.alpha, self.cutoff_upper)
        dist = dist.unsqueeze(-1)
        return self.cutoff_fn(dist) * torch.exp(
            -self.betas * (torch.exp(self.alpha * (-dist + self.cutoff_lower)) - self.means) ** 2)
