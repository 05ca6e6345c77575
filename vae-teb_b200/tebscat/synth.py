"""Synthetic CTG-shaped inputs used by bench.py and the tests.

The reference ships no data generator; these are the fixed synthetic signals the
benchmark and parity tests are quoted on: a 4 Hz fetal-heart-rate trace (baseline,
slow oscillation, random walk, noise, a few Gaussian decelerations) and a uterine
pressure trace (periodic contractions plus noise), fp32, shape (B, 2, N).
"""
import math

import torch


def ctg_batch(B: int, N: int, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(N, dtype=torch.float64) / 4.0
    ph = torch.rand(B, 1, generator=g, dtype=torch.float64) * 2 * math.pi
    ps = torch.rand(B, 1, generator=g, dtype=torch.float64) * 2 * math.pi
    fhr = 140 + 10 * torch.sin(2 * math.pi * t / 300 + ph)
    fhr = fhr + 0.15 * torch.cumsum(torch.randn(B, N, generator=g, dtype=torch.float64), dim=1)
    fhr = fhr + torch.randn(B, N, generator=g, dtype=torch.float64)
    n_dec = torch.randint(0, 4, (B,), generator=g)
    for k in range(3):
        on = (n_dec > k).to(torch.float64)[:, None]
        depth = 20 + 20 * torch.rand(B, 1, generator=g, dtype=torch.float64)
        width = 30 + 30 * torch.rand(B, 1, generator=g, dtype=torch.float64)
        centre = torch.rand(B, 1, generator=g, dtype=torch.float64) * (N / 4.0)
        fhr = fhr - on * depth * torch.exp(-0.5 * ((t - centre) / (width / 2.355)) ** 2)
    up = 15 + 35 * torch.relu(torch.sin(2 * math.pi * t / 180 + ps)) ** 4
    up = up + 2 * torch.randn(B, N, generator=g, dtype=torch.float64)
    return torch.stack([fhr, up], dim=1).to(torch.float32).contiguous()


def randn_batch(B: int, N: int, channels: int = 2, seed: int = 4321) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, channels, N, generator=g, dtype=torch.float32).contiguous()
