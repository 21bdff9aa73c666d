"""Drop-in for the part of the reference's utils.py that the inference path uses:
`sigmoid_beta_schedule` (reference utils.py:30-48).  One-off host-side float64 math, as in the reference."""
from __future__ import annotations

import torch


def sigmoid_beta_schedule(timesteps, start=-3, end=3, tau=1.0, clamp_min=1e-4):
    """Sigmoid noise schedule (arXiv:2212.11972, fig. 8) rescaled into [clamp_min, 1]; returns float64
    betas[timesteps] clipped to [0, 0.999]."""
    grid = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
    s_lo = torch.tensor(start / tau).sigmoid()
    s_hi = torch.tensor(end / tau).sigmoid()
    abar = (s_hi - ((grid * (end - start) + start) / tau).sigmoid()) / (s_hi - s_lo)
    abar = abar / abar[0]
    abar = abar * (1 - clamp_min) + clamp_min
    return torch.clip(1 - abar[1:] / abar[:-1], 0, 0.999)
