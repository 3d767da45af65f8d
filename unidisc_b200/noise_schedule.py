"""Noise schedules of the hot path (reference models/noise_schedule.py:13-157).  Only what `Diffusion` needs."""
from __future__ import annotations

import torch
import torch.nn as nn


class LogLinearNoise(nn.Module):
    """sigma(t) = -log1p(-(1-eps) t);  sigma'(t) = (1-eps) / (1 - (1-eps) t)   (noise_schedule.py:128-157)."""

    def __init__(self, eps: float = 1e-3):
        super().__init__()
        self.eps = eps
        self.sigma_max = self.total_noise(torch.tensor(1.0, dtype=torch.float32))
        self.sigma_min = self.eps + self.total_noise(torch.tensor(0.0, dtype=torch.float32))

    def rate_noise(self, t):
        return (1 - self.eps) / (1 - (1 - self.eps) * t)

    def total_noise(self, t):
        return -torch.log1p(-(1 - self.eps) * t)

    def forward(self, t):
        return self.total_noise(t), self.rate_noise(t)


def get_noise(config, dtype=torch.float32):
    ntype = getattr(config.noise, "type", "loglinear")
    if ntype == "loglinear":
        return LogLinearNoise()
    raise NotImplementedError(f"unidisc_b200: noise.type={ntype!r} is outside the hot path (only loglinear is built)")
