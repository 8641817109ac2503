"""``weighted_sdxl_mse_loss`` (train.py:2408-2416) as one fused kernel, usable from autograd.

``mean_b( mean_chw((pred.float() - target.float())**2) * table[clamp(timestep, 0, len-1)] )`` and its gradient with
respect to ``pred`` are produced in a single pass (``aoz_mse_loss``); the reference runs ~10 ATen launches.
"""
from __future__ import annotations

import torch

from . import _lib, ops


class _WeightedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, timesteps, table):
        p = pred.detach().contiguous()
        if p.dtype != torch.bfloat16:
            raise _lib.AozoraError("weighted_sdxl_mse_loss: pred must be bf16 on the B200 path")
        t = target.detach().float().contiguous()
        ts = timesteps.detach().long().contiguous()
        tb = None if table is None else table.detach().to(device=p.device, dtype=torch.float32).contiguous()
        loss, _, dpred = ops.mse_loss(p, t, ts, tb, denom=p.shape[0], grad_scale=1.0 / p.shape[0], pred_nhwc=False,
                                      need_grad=ctx.needs_input_grad[0])
        ctx.dpred = dpred
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        d = ctx.dpred
        ctx.dpred = None
        if d is None:
            return None, None, None, None
        # upstream factor (e.g. 1/GRADIENT_ACCUMULATION_STEPS from train.py:2765): exact in bf16 for powers of two
        return d * gout.to(d.dtype), None, None, None


def weighted_sdxl_mse_loss(pred, target, timesteps, timestep_loss_weights=None):
    if not pred.is_cuda:
        raise _lib.AozoraError("weighted_sdxl_mse_loss: CUDA tensors required (no CPU fallback)")
    return _WeightedMSE.apply(pred, target, timesteps, timestep_loss_weights)
