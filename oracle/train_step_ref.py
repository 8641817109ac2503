"""The reference's micro-step (train.py:2719-2784) restated around the oracle pieces (ORACLE ONLY).

``ref_train_step`` is what ``bench.py --impl reference`` and the ``cpu_baseline`` leg time
on the host cores, and what the GPU parity tests compare losses / gradients / updated
parameters against.  ``optimizer`` may be the reference's own ``RavenAdamW`` (when
/root/reference is importable) or ``RefRaven`` below (a restatement of raven.py:89-149
that travels to the GPU box).
"""
from __future__ import annotations

import contextlib

import torch

from . import host_ref


class RefRaven(torch.optim.Optimizer):
    """Per-tensor loop restatement of RavenAdamW.step (raven.py:89-149); CPU state in momentum dtype."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.98), weight_decay=0.06, eps=1e-8,
                 debias_strength=0.9, momentum_dtype=torch.bfloat16):
        super().__init__(params, dict(lr=lr, betas=betas, weight_decay=weight_decay, eps=eps,
                                      debias_strength=debias_strength, momentum_dtype=momentum_dtype))
        self._momentum_dtype = momentum_dtype

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "step" not in st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, device="cpu", dtype=group["momentum_dtype"])
                    st["exp_avg_sq"] = torch.zeros_like(p, device="cpu", dtype=group["momentum_dtype"])
                st["step"] += 1
                host_ref.raven_update_(p, p.grad, st["exp_avg"], st["exp_avg_sq"], lr=group["lr"],
                                       betas=group["betas"], eps=group["eps"], weight_decay=group["weight_decay"],
                                       debias_strength=group["debias_strength"], step=st["step"])


def make_targets(prediction_type, latents, noise, timesteps, scheduler, seed, micro_step):
    """train.py:2743-2758.  Returns (noisy_latents fp32, target fp32, conditioning timesteps)."""
    if prediction_type == "rectified_flow":
        jitter = host_ref.rf_jitter(timesteps.shape[0], seed, micro_step, device=latents.device.type)
        t = ((timesteps.float() + jitter) / 1000.0).clamp(0.0, 1.0)
        te = t.view(-1, 1, 1, 1)
        return (1 - te) * latents + te * noise, noise - latents, t * 1000.0
    noisy = scheduler.add_noise(latents, noise, timesteps)
    target = scheduler.get_velocity(latents, noise, timesteps) if prediction_type == "v_prediction" else noise
    return noisy, target, timesteps


def ref_forward_loss(unet, scheduler, batch, *, prediction_type, timesteps, micro_step, seed,
                     loss_table=None, compute_dtype=torch.bfloat16, autocast=False, taps=None):
    """Noise -> target -> UNet -> weighted MSE (train.py:2719-2763).  ``batch`` holds latents,
    embeds, pooled and time_ids_data (python lists, converted in compute_dtype as train.py:2731)."""
    latents = batch["latents"]
    ctx_mgr = torch.autocast("cpu", dtype=compute_dtype) if autocast else contextlib.nullcontext()
    with ctx_mgr:
        time_ids = torch.tensor(batch["time_ids_data"], dtype=compute_dtype)
        noise = host_ref.step_noise(latents.shape, seed, micro_step)
        noisy, target, cond = make_targets(prediction_type, latents, noise, timesteps, scheduler, seed, micro_step)
        pred = unet(noisy.to(compute_dtype), cond, batch["embeds"],
                    added_cond_kwargs={"text_embeds": batch["pooled"], "time_ids": time_ids}, taps=taps).sample
        loss = host_ref.weighted_mse(pred, target, timesteps, loss_table)
    return loss, pred, target, noisy


def ref_train_step(unet, scheduler, optimizer, batch, *, prediction_type, timesteps, micro_step, seed,
                   loss_table=None, compute_dtype=torch.bfloat16, autocast=False, grad_accum=1,
                   clip_grad_norm=1.0, do_optimizer_step=True):
    """One micro-step incl. backward, clip and optimizer step (train.py:2719-2784)."""
    loss, _, _, _ = ref_forward_loss(unet, scheduler, batch, prediction_type=prediction_type, timesteps=timesteps,
                                     micro_step=micro_step, seed=seed, loss_table=loss_table,
                                     compute_dtype=compute_dtype, autocast=autocast)
    (loss / grad_accum).backward()
    out = {"loss": float(loss.detach())}
    if do_optimizer_step:
        params = [p for g in optimizer.param_groups for p in g["params"]]
        norm = torch.nn.utils.clip_grad_norm_(params, clip_grad_norm if clip_grad_norm > 0 else float("inf"))
        out["grad_norm"] = float(norm)
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
    return out
