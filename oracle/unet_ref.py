"""Plain-torch CPU restatement of diffusers' SDXL ``UNet2DConditionModel`` (ORACLE ONLY).

PARITY UNPINNED: the arithmetic lives in ``diffusers>=0.32.0`` (reference
requirements.txt:2), which is absent from /root/reference and not installable here.
This file restates its published SDXL forward semantics ("What a CPU
restatement must follow") and is anchored on the reference's call contract
(train.py:2760-2761), on the parameter census and on the reference's own key map
(train.py:2418-2465).  Module registration order follows diffusers so that
``parameters()`` enumerates tensors in the order Raven's checkpoint indices assume
(raven.py:156-169; SURVEY.md 8b "Parameter order").

It is deliberately written with stock ``torch.nn`` layers in NCHW so that it shares no
code, layout or kernel with the product path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class RefUNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers_per_block: tuple = (1, 2, 10)
    attention_head_dim: int = 64          # every SDXL attention uses 64-wide heads
    cross_attention_dim: int = 2048
    norm_num_groups: int = 32
    addition_time_embed_dim: int = 256
    pooled_dim: int = 1280                # projection_class_embeddings_input_dim = pooled + 6*256 = 2816
    down_has_attn: tuple = (False, True, True)

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4

    @property
    def add_in_dim(self):
        return self.pooled_dim + 6 * self.addition_time_embed_dim


def sdxl_config():
    return RefUNetConfig()


def tiny_config():
    """Same topology as SDXL (3 levels, cross-attn on levels 1-2, skip concat), small widths."""
    return RefUNetConfig(block_out_channels=(64, 128, 256), transformer_layers_per_block=(1, 1, 2),
                         cross_attention_dim=128, addition_time_embed_dim=32, pooled_dim=64)


def timestep_embedding(timesteps, dim, max_period=10000.0):
    """diffusers ``Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)`` [3P]: cat(cos, sin)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    ang = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)
        else:
            self.conv_shortcut = None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, ctx_dim, head_dim):
        super().__init__()
        self.heads = dim // head_dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, T, C = x.shape
        h = self.heads
        q = self.to_q(x).view(B, T, h, C // h).transpose(1, 2)
        k = self.to_k(ctx).view(B, ctx.shape[1], h, C // h).transpose(1, 2)
        v = self.to_v(ctx).view(B, ctx.shape[1], h, C // h).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)           # no mask, no dropout, scale 1/sqrt(64)
        o = o.transpose(1, 2).reshape(B, T, C)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)                               # exact erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, ctx_dim, head_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, dim, head_dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, ctx_dim, head_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, depth, ctx_dim, head_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(dim, ctx_dim, head_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x).permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        h = self.proj_out(h)
        return h.reshape(B, H, W, C).permute(0, 3, 1, 2) + res


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    """DownBlock2D / CrossAttnDownBlock2D; ``attentions`` registered before ``resnets`` [3P]."""

    def __init__(self, cfg, cin, cout, depth, has_attn, add_down):
        super().__init__()
        if has_attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, depth, cfg.cross_attention_dim, cfg.attention_head_dim, cfg.norm_num_groups)
                for _ in range(cfg.layers_per_block)])
        else:
            self.attentions = None
        self.resnets = nn.ModuleList([
            ResnetBlock2D(cin if i == 0 else cout, cout, cfg.time_embed_dim, cfg.norm_num_groups)
            for i in range(cfg.layers_per_block)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg, c, depth):
        super().__init__()
        self.attentions = nn.ModuleList([
            Transformer2DModel(c, depth, cfg.cross_attention_dim, cfg.attention_head_dim, cfg.norm_num_groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, cfg.time_embed_dim, cfg.norm_num_groups) for _ in range(2)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg, cin, cout, cprev, depth, has_attn, add_up):
        super().__init__()
        n = cfg.layers_per_block + 1
        if has_attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, depth, cfg.cross_attention_dim, cfg.attention_head_dim, cfg.norm_num_groups)
                for _ in range(n)])
        else:
            self.attentions = None
        resnets = []
        for i in range(n):
            skip = cin if i == n - 1 else cout
            rin = cprev if i == 0 else cout
            resnets.append(ResnetBlock2D(rin + skip, cout, cfg.time_embed_dim, cfg.norm_num_groups))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, ctx):
        for i, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class RefUNet2DConditionModel(nn.Module):
    """Oracle restatement of the SDXL UNet; ``forward`` mirrors train.py:2760-2761's call."""

    def __init__(self, cfg: RefUNetConfig | None = None):
        super().__init__()
        cfg = cfg or sdxl_config()
        self.cfg = cfg
        self.config = SimpleNamespace(in_channels=cfg.in_channels, out_channels=cfg.out_channels)
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], cfg.time_embed_dim)
        self.add_embedding = TimestepEmbedding(cfg.add_in_dim, cfg.time_embed_dim)
        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()          # registered before mid_block, as in diffusers [3P]
        cout = boc[0]
        for i, c in enumerate(boc):
            cin, cout = cout, c
            self.down_blocks.append(DownBlock(cfg, cin, cout, cfg.transformer_layers_per_block[i],
                                              cfg.down_has_attn[i], add_down=(i != len(boc) - 1)))
        self.mid_block = MidBlock(cfg, boc[-1], cfg.transformer_layers_per_block[-1])
        rev = list(reversed(boc))
        rdepth = list(reversed(cfg.transformer_layers_per_block))
        rattn = list(reversed(cfg.down_has_attn))
        cout = rev[0]
        for i, c in enumerate(rev):
            cprev, cout = cout, c
            cin = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(cfg, cin, cout, cprev, rdepth[i], rattn[i], add_up=(i != len(boc) - 1)))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def embed(self, timestep, text_embeds, time_ids, dtype):
        cfg = self.cfg
        B = text_embeds.shape[0]
        if timestep.dim() == 0:
            timestep = timestep.expand(B)
        t_emb = timestep_embedding(timestep, cfg.block_out_channels[0]).to(dtype)
        emb = self.time_embedding(t_emb)
        tid = timestep_embedding(time_ids.flatten(), cfg.addition_time_embed_dim).reshape(B, -1)
        add = torch.cat([text_embeds, tid.to(text_embeds.dtype)], dim=-1).to(dtype)
        return emb + self.add_embedding(add)

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None, taps=None):
        dtype = self.conv_in.weight.dtype
        emb = self.embed(timestep, added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"], dtype)
        ctx = encoder_hidden_states.to(dtype)
        x = self.conv_in(sample.to(dtype))
        skips = [x]
        for bi, blk in enumerate(self.down_blocks):
            x, outs = blk(x, emb, ctx)
            skips.extend(outs)
            if taps is not None:
                taps[f"down_blocks.{bi}"] = x
        x = self.mid_block(x, emb, ctx)
        if taps is not None:
            taps["mid_block"] = x
        for bi, blk in enumerate(self.up_blocks):
            x = blk(x, skips, emb, ctx)
            if taps is not None:
                taps[f"up_blocks.{bi}"] = x
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return SimpleNamespace(sample=x)

    # reference boundary no-ops (train.py:204-228, 2660)
    def enable_gradient_checkpointing(self):
        pass

    def set_attn_processor(self, proc):
        pass


def init_weights_(model: nn.Module, seed: int = 42, std: float = 0.02):
    """Fixed synthetic init shared by oracle and product: N(0, std) for >=2-D
    weights, small N(0, std) biases, norm affine 1 + N(0, std) / N(0, std).  Drawn in
    ``named_parameters()`` order from one CPU generator so any module with the same names,
    shapes and order receives identical tensors."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            t = torch.randn(p.shape, generator=g, dtype=torch.float32) * std
            if p.dim() == 1 and name.endswith("weight"):
                t = t + 1.0
            p.copy_(t.to(p.dtype))
    return model
