"""Checkpoint I/O around the training step: the SDXL single-file layout, model export and the
training-state file, in the reference's formats.

* ``ldm_key`` / ``unet_key_mapping``: diffusers UNet key -> LDM ("model.diffusion_model.*") key, computed structurally from
  the parsed name (block index arithmetic of the SDXL layout) -- the mapping the reference builds with replacement tables
  (train.py:2418-2465).  ``tests/test_checkpoint.py`` checks all 1680 keys against the reference's own function.
* ``load_unet_single_file`` (= ``UNet2DConditionModel.from_single_file``, train.py:1439-1470): reads only the UNet tensors of a
  ``.safetensors`` single-file checkpoint through the INVERSE mapping, peeks in/out channel counts like the reference.
* ``save_model`` (train.py:2467-2513): base checkpoint tensors cast to the compute dtype, UNet tensors merged under their LDM
  names, written as one ``.safetensors``.
* ``save_training_state`` / ``load_training_state`` (train.py:2515-2531, 2566-2582): the ``.pt`` resume file with the Raven
  moments in ``save_cpu_state`` format (gathered across ranks under data parallel), sampler position and RNG states.
"""
from __future__ import annotations

import random
import re
from pathlib import Path

import numpy as np
import torch

LDM_PREFIX = "model.diffusion_model."

_TOP = {"time_embedding.linear_1": "time_embed.0", "time_embedding.linear_2": "time_embed.2", "conv_in": "input_blocks.0.0",
        "conv_norm_out": "out.0", "conv_out": "out.2", "add_embedding.linear_1": "label_emb.0.0", "add_embedding.linear_2": "label_emb.0.2"}
_RESNET = {"norm1": "in_layers.0", "conv1": "in_layers.2", "norm2": "out_layers.0", "conv2": "out_layers.3",
           "time_emb_proj": "emb_layers.1", "conv_shortcut": "skip_connection"}
_BLOCK = re.compile(r"^(down_blocks|up_blocks)\.(\d+)\.(resnets|attentions|downsamplers|upsamplers)\.(\d+)\.(.*)$")
_MID = re.compile(r"^mid_block\.(resnets|attentions)\.(\d+)\.(.*)$")


def _resnet_tail(tail: str) -> str:
    head, _, rest = tail.partition(".")
    return f"{_RESNET.get(head, head)}.{rest}" if rest else _RESNET.get(head, head)


def ldm_key(hf_key: str) -> str:
    """LDM name of one diffusers SDXL-UNet parameter.  Layout facts used: three levels, two resnets per down level and
    three per up level, ``input_blocks[3*i + j + 1]`` / ``output_blocks[3*i + j]`` hold (resnet, transformer) pairs, a
    level's downsampler is ``input_blocks[3*(i+1)].0.op``, its upsampler the last entry of ``output_blocks[3*i + 2]``."""
    for hf, ldm in _TOP.items():
        if hf_key.startswith(hf + "."):
            return LDM_PREFIX + ldm + hf_key[len(hf):]
    m = _MID.match(hf_key)
    if m:
        kind, j, tail = m.group(1), int(m.group(2)), m.group(3)
        if kind == "attentions":
            return f"{LDM_PREFIX}middle_block.1.{tail}"
        return f"{LDM_PREFIX}middle_block.{2 * j}.{_resnet_tail(tail)}"
    m = _BLOCK.match(hf_key)
    if not m:
        return hf_key if hf_key.startswith(LDM_PREFIX) else LDM_PREFIX + hf_key
    side, i, kind, j, tail = m.group(1), int(m.group(2)), m.group(3), int(m.group(4)), m.group(5)
    if side == "down_blocks":
        if kind == "resnets":
            return f"{LDM_PREFIX}input_blocks.{3 * i + j + 1}.0.{_resnet_tail(tail)}"
        if kind == "attentions":
            return f"{LDM_PREFIX}input_blocks.{3 * i + j + 1}.1.{tail}"
        return f"{LDM_PREFIX}input_blocks.{3 * (i + 1)}.0.op.{tail.split('.', 1)[1]}"          # downsamplers.0.conv.* -> op.*
    if kind == "resnets":
        return f"{LDM_PREFIX}output_blocks.{3 * i + j}.0.{_resnet_tail(tail)}"
    if kind == "attentions":
        return f"{LDM_PREFIX}output_blocks.{3 * i + j}.1.{tail}"
    # upsampler: after (resnet, transformer) on attention levels, directly after the resnet on the last level
    return f"{LDM_PREFIX}output_blocks.{3 * i + 2}.{2 if i < 2 else 1}.{tail}"


def unet_key_mapping(hf_keys):
    """{diffusers key: LDM key} in the given order (the reference's ``get_unet_key_mapping``)."""
    return {k: ldm_key(k) for k in hf_keys}


def peek_unet_channels(path):
    """(in_channels, out_channels) of a single-file checkpoint without loading it (train.py:1439-1455)."""
    from safetensors import safe_open
    cin = cout = 4
    with safe_open(str(path), framework="pt", device="cpu") as f:
        keys = set(f.keys())
        if LDM_PREFIX + "input_blocks.0.0.weight" in keys:
            cin = f.get_slice(LDM_PREFIX + "input_blocks.0.0.weight").get_shape()[1]
        if LDM_PREFIX + "out.2.weight" in keys:
            cout = f.get_slice(LDM_PREFIX + "out.2.weight").get_shape()[0]
    return cin, cout


def load_unet_single_file(path, torch_dtype=torch.bfloat16, device="cpu", in_channels=None, out_channels=None, config=None, strict=True):
    """Build the B200 ``UNet2DConditionModel`` from an SDXL single-file ``.safetensors`` (only UNet tensors are read)."""
    from safetensors import safe_open

    from .unet import UNet2DConditionModel, sdxl_config
    if in_channels is None or out_channels is None:
        pin, pout = peek_unet_channels(path)
        in_channels = pin if in_channels is None else in_channels
        out_channels = pout if out_channels is None else out_channels
    cfg = config or sdxl_config(in_channels=in_channels, out_channels=out_channels)
    with torch.device("meta"):
        model = UNet2DConditionModel(cfg)
    model = model.to_empty(device=device).to(torch_dtype)
    want = unet_key_mapping(list(model.state_dict().keys()))
    missing = []
    with safe_open(str(path), framework="pt", device="cpu") as f:
        have = set(f.keys())
        sd = model.state_dict()
        with torch.no_grad():
            for hf, ldm in want.items():
                if ldm not in have:
                    missing.append(ldm)
                    continue
                t = f.get_tensor(ldm)
                if tuple(t.shape) != tuple(sd[hf].shape):
                    raise ValueError(f"{ldm}: checkpoint shape {tuple(t.shape)} != model shape {tuple(sd[hf].shape)} ({hf})")
                sd[hf].copy_(t.to(torch_dtype))
    if missing and strict:
        raise KeyError(f"{len(missing)} UNet tensors missing from {path}, e.g. {missing[:3]}")
    return model


def save_model(output_path, unet, base_checkpoint_path, compute_dtype):
    """Write a full single-file checkpoint: every float tensor of the base file cast to ``compute_dtype``, the trained UNet
    tensors replacing (or adding) their LDM-named entries (train.py:2467-2513).  Parameters that are views of a shared buffer
    (stacked q/k/v, data-parallel flat storage) are copied out individually, so the file never aliases memory."""
    from safetensors.torch import load_file, save_file
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    tensors = load_file(str(base_checkpoint_path), device="cpu")
    for k, t in tensors.items():
        if t.dtype in (torch.float32, torch.float16, torch.bfloat16):
            tensors[k] = t.to(dtype=compute_dtype)
    state = unet.state_dict()
    added = []
    for hf, ldm in unet_key_mapping(list(state.keys())).items():
        if ldm not in tensors:
            added.append(ldm)
        tensors[ldm] = state[hf].detach().to("cpu", dtype=compute_dtype).contiguous().clone()
    save_file(tensors, str(output_path))
    return added


def save_training_state(path, *, global_step, micro_step, optimizer, sampler_seed, sampler_epoch, timestep_sampler=None):
    """The reference's ``*_training_state_step_N.pt`` (train.py:2522-2531): Raven/Titan moments in ``save_cpu_state`` format
    (index-keyed, CPU tensors; gathered from every rank's shard under data parallel), schedule position, RNG states."""
    optim_state = optimizer.save_cpu_state() if hasattr(optimizer, "save_cpu_state") else optimizer.state_dict()
    state = {"global_step": global_step, "micro_step": micro_step, "optimizer_state": optim_state, "sampler_seed": sampler_seed,
             "sampler_epoch": max(int(sampler_epoch) - 1, 0),
             "timestep_sampler_state": timestep_sampler.state_dict() if timestep_sampler is not None and hasattr(timestep_sampler, "state_dict") else None,
             "random_state": random.getstate(), "numpy_state": np.random.get_state(), "torch_cpu_state": torch.get_rng_state(),
             "torch_cuda_state": torch.cuda.get_rng_state() if torch.cuda.is_available() else None}
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    torch.save(state, path)
    return state


def load_training_state(path, *, optimizer=None, timestep_sampler=None, grad_accum=1, restore_rng=True):
    """Inverse of ``save_training_state`` with the reference's resume rules (train.py:2566-2582): ``micro_step`` falls back to
    ``global_step * GRADIENT_ACCUMULATION_STEPS``; RNG states are restored when present.  Returns the loaded dictionary with
    ``micro_step`` / ``optimizer_step`` filled in."""
    state = torch.load(Path(path), map_location="cpu", weights_only=False)
    gs = state.get("global_step", 0)
    state["micro_step"] = state.get("micro_step", gs * grad_accum)
    state["optimizer_step"] = state["micro_step"] // max(1, grad_accum)
    if optimizer is not None and state.get("optimizer_state") is not None:
        if hasattr(optimizer, "load_cpu_state"):
            optimizer.load_cpu_state(state["optimizer_state"])
        else:
            optimizer.load_state_dict(state["optimizer_state"])
    if timestep_sampler is not None and state.get("timestep_sampler_state") is not None and hasattr(timestep_sampler, "load_state_dict"):
        timestep_sampler.load_state_dict(state["timestep_sampler_state"])
    if restore_rng:
        if "random_state" in state:
            random.setstate(state["random_state"])
        if "numpy_state" in state:
            np.random.set_state(state["numpy_state"])
        if "torch_cpu_state" in state:
            torch.set_rng_state(state["torch_cpu_state"])
        if state.get("torch_cuda_state") is not None and torch.cuda.is_available():
            torch.cuda.set_rng_state(state["torch_cuda_state"])
    return state
