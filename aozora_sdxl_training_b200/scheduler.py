"""``DDPMScheduler`` as the reference uses it (train.py:2613-2619, 2626, 2755-2757): SDXL-base scheduler config
(scaled_linear betas in [0.00085, 0.012], 1000 steps, no zero-terminal-SNR), ``alphas_cumprod``, ``add_noise`` and
``get_velocity``.  diffusers is third-party and absent from the reference tree; these follow its published
semantics (``alphas_cumprod`` is cast to the sample dtype before the square roots).  The fused training step does not
call ``add_noise`` / ``get_velocity`` -- it uses ``ops.noise_target`` (one kernel) -- they exist for drop-in callers.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 prediction_type="epsilon"):
        if beta_schedule != "scaled_linear":
            raise ValueError("only the SDXL 'scaled_linear' schedule is implemented")
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.config = SimpleNamespace(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps,
                                      beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule)

    @classmethod
    def from_pretrained(cls, repo=None, subfolder=None, **kwargs):
        """The reference loads stabilityai/stable-diffusion-xl-base-1.0 'scheduler' from the hub; its values are the
        defaults above, so no network access is needed."""
        return cls()

    def _coeffs(self, ref, timesteps):
        acp = self.alphas_cumprod.to(device=ref.device).to(dtype=ref.dtype)
        t = timesteps.to(ref.device)
        a = (acp[t] ** 0.5).flatten()
        b = ((1 - acp[t]) ** 0.5).flatten()
        while a.dim() < ref.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a, b

    def add_noise(self, original_samples, noise, timesteps):
        a, b = self._coeffs(original_samples, timesteps)
        return a * original_samples + b * noise

    def get_velocity(self, sample, noise, timesteps):
        a, b = self._coeffs(sample, timesteps)
        return a * noise - b * sample
