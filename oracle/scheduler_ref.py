"""Restatement of diffusers ``DDPMScheduler`` as train.py uses it (ORACLE ONLY).

PARITY UNPINNED (third-party ``diffusers>=0.32.0``, absent).  Follows the SDXL-base
scheduler config the reference loads at train.py:2613-2616: ``scaled_linear`` betas in
[0.00085, 0.012], 1000 steps, no zero-terminal-SNR rescale; ``add_noise`` /
``get_velocity`` cast ``alphas_cumprod`` to the sample dtype before the square roots
(SURVEY.md row a3), which matters because the cached latents are bf16.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch


class RefDDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, prediction_type="epsilon"):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.config = SimpleNamespace(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps)

    def _coeffs(self, ref, timesteps):
        acp = self.alphas_cumprod.to(device=ref.device).to(dtype=ref.dtype)
        t = timesteps.to(ref.device)
        a = (acp[t] ** 0.5).flatten()
        b = ((1 - acp[t]) ** 0.5).flatten()
        while a.dim() < ref.dim():
            a = a.unsqueeze(-1)
            b = b.unsqueeze(-1)
        return a, b

    def add_noise(self, original_samples, noise, timesteps):
        a, b = self._coeffs(original_samples, timesteps)
        return a * original_samples + b * noise

    def get_velocity(self, sample, noise, timesteps):
        a, b = self._coeffs(sample, timesteps)
        return a * noise - b * sample
