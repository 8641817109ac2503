"""The reference's training loop (train.py:2558-2828 ``main``) over the B200 pieces: cached dataset -> bucketed batch
schedule -> pinned-memory feeder -> ``SDXLTrainStep`` (noise, UNet, loss, reverse sweep, clip, Raven) -> periodic export.

Only what the loop needs is here (no GUI reporter; the caching pass is ``ensure_cache`` with the caller's encoders): it exists so that a run can be started,
checkpointed and resumed with the reference's files -- same cache directory, same ``.safetensors`` / ``.pt`` outputs, same
schedule and ticket positions after a resume (train.py:2566-2582, 2700-2712).
"""
from __future__ import annotations

from pathlib import Path

import torch

from . import cache_builder, checkpoint, data, host
from .optimizers import RavenAdamW
from .trainer import SDXLTrainStep


# training_utils/config/config.py:122-128 (RAVEN_PARAMS) and the TITAN_PARAMS block beside it: what ``create_optimizer`` merges under
# the run's own dictionary before reading any key (train.py:2256-2270)
_DEFAULT_OPTIMIZER_PARAMS = {"betas": [0.9, 0.999], "eps": 1e-8, "weight_decay": 0.01, "debias_strength": 0.3, "momentum_dtype": "bfloat16"}


def optimizer_kwargs(config, kind="raven"):
    """Constructor keywords exactly as ``create_optimizer`` forms them (train.py:2256-2270): package defaults overlaid by the
    config's ``RAVEN_PARAMS`` / ``TITAN_PARAMS``; ``momentum_dtype`` arrives as a string ("bfloat16" -> bf16, anything else -> fp32);
    ``betas`` as a list; keys absent from both fall back to eps 1e-8, weight_decay 0.01, debias_strength 1.0; lr = the LR curve's peak."""
    curve = getattr(config, "LR_CUSTOM_CURVE", None) or []
    lr = max(pt[1] for pt in curve) if curve else getattr(config, "LEARNING_RATE", 1e-6)
    merged = {**_DEFAULT_OPTIMIZER_PARAMS, **(getattr(config, f"{kind.upper()}_PARAMS", None) or {})}
    md = merged.get("momentum_dtype", "bfloat16")
    if not isinstance(md, torch.dtype):
        md = torch.bfloat16 if md == "bfloat16" else torch.float32
    return dict(lr=lr, betas=tuple(merged.get("betas", [0.9, 0.999])), eps=merged.get("eps", 1e-8),
                weight_decay=merged.get("weight_decay", 0.01), debias_strength=merged.get("debias_strength", 1.0), momentum_dtype=md)


def build_optimizer(config, unet, dp=None):
    """The reference's optimizer construction (train.py:2664-2679, 2256-2270): exclusion keywords applied, ONE parameter group in
    ``unet.parameters()`` order with ``lr_scale`` 1.0, RavenAdamW or -- ``OPTIMIZER_TYPE == "titan"`` -- TitanAdamW."""
    host.apply_exclusion(unet, list(getattr(config, "UNET_EXCLUDE_TARGETS", []) or []))
    kind = str(getattr(config, "OPTIMIZER_TYPE", "raven")).lower()
    if kind not in ("raven", "titan"):
        raise ValueError(f"Unsupported optimizer type: '{getattr(config, 'OPTIMIZER_TYPE', None)}'")
    kw = optimizer_kwargs(config, kind)
    if dp is not None:
        if kind != "raven":
            raise ValueError("data-parallel training shards Raven state; TitanAdamW's gradient offload has no sharded form")
        dp.momentum_dtype = kw.pop("momentum_dtype")           # the shards are allocated by make_optimizer, in the config's dtype
        return dp.make_optimizer(**kw)
    group = [{"params": [p for p in unet.parameters() if p.requires_grad], "lr_scale": 1.0}]
    if kind == "titan":
        from .optimizers import TitanAdamW
        return TitanAdamW(group, **kw)
    return RavenAdamW(group, **kw)


def ensure_cache(config, encoders, device="cuda", dp=None) -> bool:
    """The caching pass in front of the loop (train.py:2575-2602): build / refresh the latent and text cache when it is missing or
    stale.  ``encoders`` = (tokenizer_1, tokenizer_2, text_encoder_1, text_encoder_2, vae).  Under data parallel rank 0 builds
    while the other ranks wait at a barrier.  Returns True if a build ran."""
    if not hasattr(config, "is_rectified_flow"):
        config.is_rectified_flow = getattr(config, "PREDICTION_TYPE", "epsilon") == "rectified_flow"
    ran = False
    if dp is None or dp.rank == 0:
        if cache_builder.cache_needs_build(config):
            cache_builder.build_cache(config, *encoders, device)
            ran = True
    if dp is not None:
        import torch.distributed as dist
        dist.barrier(group=dp.group)
    return ran


def run_training(config, unet, *, device="cuda", dp=None, optimizer=None, resume_state_path=None, base_checkpoint_path=None,
                 use_cuda_graph=False, on_step=None):
    """Train for ``config.MAX_TRAIN_STEPS`` micro-steps (or resume).  Returns {"losses", "micro_step", "saved"}.

    ``config`` carries the reference's flat keys: SEED, BATCH_SIZE (global), MAX_TRAIN_STEPS, GRADIENT_ACCUMULATION_STEPS,
    INSTANCE_DATASETS, PREDICTION_TYPE, LR_CUSTOM_CURVE, CLIP_GRAD_NORM, SAVE_EVERY_N_STEPS, OUTPUT_DIR, OUTPUT_NAME ...
    ``base_checkpoint_path``: the single-file checkpoint whose non-UNet tensors an export carries along; without it only the
    training-state file is written."""
    if not hasattr(config, "is_rectified_flow"):
        config.is_rectified_flow = getattr(config, "PREDICTION_TYPE", "epsilon") == "rectified_flow"
    world = 1 if dp is None else dp.world
    rank = 0 if dp is None else dp.rank
    optimizer = optimizer or build_optimizer(config, unet, dp)
    step = SDXLTrainStep(unet, optimizer, config, device=device, dp=dp, use_cuda_graph=use_cuda_graph)
    dataset = data.CachedLatentDataset(config)
    start, sampler_seed = 0, step.seed
    if resume_state_path is not None:
        st = checkpoint.load_training_state(resume_state_path, optimizer=optimizer, timestep_sampler=step.sampler,
                                            grad_accum=step.grad_accum)
        start = int(st["micro_step"])
        step.micro_step = start
        step.optimizer_steps = int(st["optimizer_step"])
        sampler_seed = int(st.get("sampler_seed", sampler_seed))             # train.py:2564: the schedule keeps the saved seed
        if st.get("timestep_sampler_state") is None and start > 0:            # train.py:2644-2645: older state files
            step.sampler.set_current_step(start)
        if step.lr_scheduler is not None:
            step.lr_scheduler.step(start)                                     # train.py:2687
    # train.py:2651-2660: epoch-shuffled buckets, or -- TIMESTEP_FORCE_IMAGE_BIN_SPREAD -- images spread over the ticket bins
    schedule = data.pack_sample_schedule(
        data.image_batch_schedule(dataset.bucket_keys, int(config.MAX_TRAIN_STEPS), int(config.BATCH_SIZE), sampler_seed,
                                  step.sampler.ticket_pool, step.sampler.bin_ranges,
                                  bool(getattr(config, "TIMESTEP_FORCE_IMAGE_BIN_SPREAD", False))),
        int(config.BATCH_SIZE))
    feeder = data.BatchFeeder(dataset, schedule, rank=rank, world=world, start_step=start)
    save_every = int(getattr(config, "SAVE_EVERY_N_STEPS", 0) or 0)
    out_dir = Path(getattr(config, "OUTPUT_DIR", "."))
    stem = getattr(config, "OUTPUT_NAME", "aozora_b200")
    losses, saved = [], []
    for batch in feeder:
        if step.micro_step >= int(config.MAX_TRAIN_STEPS):
            break
        if not batch:                              # every item of the batch failed to load: the reference skips it too
            continue
        res = step.step(batch)
        losses.append(res.loss)                    # one device scalar per step (read back once, after the loop)
        if on_step is not None:
            on_step(step.micro_step, res)
        if save_every and res.did_optimizer_step and step.optimizer_steps % save_every == 0:
            gs = step.optimizer_steps
            if dp is not None:
                dp.gather_params()                 # deferred all-gather: the export below reads every rank's full parameters
            state_path = out_dir / f"{stem}_training_state_step_{gs}.pt"
            # under data parallel every rank takes part in the gather inside save_cpu_state; rank 0 writes
            st = checkpoint.save_training_state(state_path if rank == 0 else out_dir / f".rank{rank}_{stem}_state.pt", global_step=gs,
                                                micro_step=step.micro_step, optimizer=optimizer, sampler_seed=sampler_seed,
                                                sampler_epoch=step.micro_step, timestep_sampler=step.sampler)
            del st
            if rank == 0:
                if base_checkpoint_path is not None:
                    checkpoint.save_model(out_dir / f"{stem}_step_{gs}.safetensors", unet, base_checkpoint_path,
                                          getattr(config, "compute_dtype", torch.bfloat16))
                saved.append(state_path)
    if dp is not None:
        dp.gather_params()
    loss_values = torch.cat([l.reshape(1).float() for l in losses]).cpu().tolist() if losses else []
    return dict(losses=loss_values, micro_step=step.micro_step, saved=saved, step=step)
