"""CPU restatement of the reference's host-side pieces of the training step (ORACLE ONLY).

PINNED: every function here is checked in ``tests/test_oracle_pinning.py`` against the
frozen outputs of the reference's own code (``tests/golden/host_golden.json``, produced by
``tests/golden/make_golden.py`` through ``oracle/ref_shim.py``) and, when /root/reference is
present, against the live reference functions.

Citations are into /root/reference.
"""
from __future__ import annotations

import fnmatch
import math
import random

import numpy as np
import torch

U64 = (1 << 64) - 1


# ----------------------------------------------------------------------------------------
# timestep tickets (train.py:577-685, 2163-2208)
# ----------------------------------------------------------------------------------------
def largest_remainder_scale(counts, target_total):
    """train.py:577-594 -- rescale integer counts to ``target_total`` by largest remainder."""
    target_total = max(0, int(target_total))
    counts = [max(0, int(c or 0)) for c in counts]
    tot = sum(counts)
    if target_total <= 0 or tot <= 0:
        return [0] * len(counts)
    exact = [c / tot * target_total for c in counts]
    base = [int(e) for e in exact]
    short = target_total - sum(base)
    if short > 0:
        # python's sort is stable and reverse=True keeps original order among ties
        by_frac = sorted(range(len(exact)), key=lambda i: exact[i] - base[i], reverse=True)
        for i in by_frac[:short]:
            base[i] += 1
    return base


def bin_counts_and_ranges(allocation, total, n_timesteps=1000):
    """train.py:597-620."""
    usable = bool(allocation) and "counts" in allocation and "bin_size" in allocation and sum(allocation["counts"]) != 0
    if not usable:
        bin_size = 100
        nb = 10
        counts = [total // nb + (1 if i < total % nb else 0) for i in range(nb)]
    else:
        bin_size = max(1, int(allocation["bin_size"]))
        counts = largest_remainder_scale(allocation["counts"], total)
    scale = n_timesteps / 1000.0
    out_c, out_r = [], []
    for i, c in enumerate(counts):
        if c <= 0:
            continue
        lo = int(i * bin_size * scale)
        hi = min(n_timesteps, max(lo + 1, int((i + 1) * bin_size * scale)))
        if lo >= n_timesteps:
            break
        out_c.append(int(c))
        out_r.append((lo, hi))
    return out_c, out_r


def balanced_bin_order(bin_counts, seed):
    """train.py:623-642 -- jittered, balanced interleave of bins (PCG64(seed+7919), lexsort)."""
    if not bin_counts:
        return []
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    pos, ids, jit = [], [], []
    for b, c in enumerate(bin_counts):
        if c <= 0:
            continue
        pos.append((np.arange(c, dtype=np.float64) + rng.random(c)) / c)
        ids.append(np.full(c, b, dtype=np.int32))
        jit.append(rng.random(c))
    if not pos:
        return []
    pos, ids, jit = np.concatenate(pos), np.concatenate(ids), np.concatenate(jit)
    return ids[np.lexsort((jit, pos))].tolist()


def stratified_pool(bin_counts, bin_ranges, seed):
    """train.py:645-662 -- per-bin no-repeat decks dealt in ``balanced_bin_order``."""
    rng = np.random.Generator(np.random.PCG64(seed))
    decks = []
    for c, (lo, hi) in zip(bin_counts, bin_ranges):
        vals = np.arange(lo, hi, dtype=np.int64)
        deck = []
        while len(deck) < c:
            deck.extend(rng.permutation(vals).tolist()[: c - len(deck)])
        decks.append(deck)
    cursor = [0] * len(decks)
    pool = []
    for b in balanced_bin_order(bin_counts, seed):
        pool.append(int(decks[b][cursor[b]]))
        cursor[b] += 1
    return pool


def ticket_pool(allocation, total, n_timesteps=1000, seed=42, stratified=False):
    """train.py:665-685.  Returns (pool, bin_ranges)."""
    total = max(0, int(total))
    n_timesteps = max(1, int(n_timesteps))
    seed = int(seed if seed else 42)
    counts, ranges = bin_counts_and_ranges(allocation, total, n_timesteps)
    if stratified:
        pool = stratified_pool(counts, ranges, seed)
    else:
        rng = np.random.Generator(np.random.PCG64(seed))
        pool = []
        for c, (lo, hi) in zip(counts, ranges):
            pool.extend(rng.integers(lo, hi, size=max(1, int(c))).tolist())
        random.Random(seed).shuffle(pool)
    if not pool:
        fb = random.Random(seed)
        pool = [fb.randint(0, n_timesteps - 1) for _ in range(total)]
    while len(pool) < total:
        pool.extend(pool[: total - len(pool)])
    return pool[:total], ranges


class RefTimestepSampler:
    """train.py:2163-2208 (pool of MAX_TRAIN_STEPS*BATCH_SIZE tickets popped sequentially)."""

    def __init__(self, max_train_steps, batch_size, seed=42, allocation=None, stratified=False):
        self.batch_size = batch_size
        self.pool, self.bin_ranges = ticket_pool(allocation, max_train_steps * batch_size, 1000,
                                                 seed if seed else 42, stratified)
        self.pool_index = 0

    def set_current_step(self, micro_step):
        self.pool_index = (micro_step * self.batch_size) % len(self.pool)

    def sample(self, n):
        out = []
        for _ in range(n):
            if self.pool_index >= len(self.pool):
                self.pool_index = 0
            out.append(self.pool[self.pool_index])
            self.pool_index += 1
        return torch.tensor(out, dtype=torch.long), out[0]


def logit_normal_counts(mu, sigma, total, bin_size=100):
    """GUI recipe for TIMESTEP_ALLOCATION.counts in Logit-Normal mode
    (gui/gui.py:5594-5603 weights, gui/gui.py:2309-2318 + gui_math.py:30-46 largest remainder)."""
    nb = math.ceil(1000 / bin_size)
    phi = lambda z: 0.5 * (1.0 + math.erf(z / math.sqrt(2.0)))
    logit = lambda p: math.log(p / (1.0 - p))
    w = []
    for i in range(nb):
        ts, te = i * bin_size, min(1000, (i + 1) * bin_size)
        a = logit(max(ts / 1000.0, 1e-6))
        b = logit(min(te / 1000.0, 1.0 - 1e-6))
        w.append(max(0.0, phi((b - mu) / sigma) - phi((a - mu) / sigma)))
    s = sum(w)
    exact = [x / s * total for x in w]
    base = [int(math.floor(e)) for e in exact]
    short = total - sum(base)
    order = sorted(range(nb), key=lambda i: exact[i] - base[i], reverse=True)
    for i in order[:short]:
        base[i] += 1
    return base


# ----------------------------------------------------------------------------------------
# generators (train.py:248-263)
# ----------------------------------------------------------------------------------------
def step_noise(shape, seed, step, device="cpu"):
    """train.py:248-254: reseed with (seed+step) % (2**32-1), draw fp32 normal."""
    g = torch.Generator(device=device)
    g.manual_seed((seed + step) % (2 ** 32 - 1))
    return torch.randn(shape, device=device, dtype=torch.float32, generator=g)


def lcg_mixed_seed(seed, *parts):
    """train.py:257-263: 64-bit LCG mix, reduced mod 2**63-1."""
    v = int(seed if seed else 42) & U64
    for p in parts:
        v = (v * 6364136223846793005 + int(p) + 1442695040888963407) & U64
    return v % (2 ** 63 - 1)


def rf_jitter(n, seed, step, device="cpu"):
    g = torch.Generator(device=device)
    g.manual_seed(lcg_mixed_seed(seed, step, 0x5D1))
    return torch.rand((n,), device=device, dtype=torch.float32, generator=g)


# ----------------------------------------------------------------------------------------
# LR curve (train.py:325-359) and loss-weight table (train.py:2351-2405)
# ----------------------------------------------------------------------------------------
def lr_at(curve_points, micro_step, total_micro_steps):
    pts = sorted([list(p) for p in curve_points], key=lambda p: p[0])
    if pts[0][0] != 0.0:
        pts.insert(0, [0.0, pts[0][1]])
    if pts[-1][0] != 1.0:
        pts.append([1.0, pts[-1][1]])
    total = max(total_micro_steps, 1)
    x = max(0.0, min(1.0, micro_step / max(total - 1, 1)))
    for (x1, y1), (x2, y2) in zip(pts[:-1], pts[1:]):
        if x1 <= x <= x2:
            if x2 - x1 == 0:
                return y1
            return y1 + (x - x1) / (x2 - x1) * (y2 - y1)
    return pts[-1][1]


def bell_table(steps=1000):
    grid = torch.arange(steps, dtype=torch.float32)
    y = torch.exp(-2.0 * ((grid - steps / 2) / steps).pow(2))
    lo = y.min()
    return (y - lo).clamp_min(0.0) * (steps / (y - lo).sum().clamp_min(1e-12))


def loss_weight_table(points, steps=1000):
    if steps <= 0:
        return torch.ones(1)
    if not points:
        return torch.ones(steps)
    if isinstance(points, dict):
        return bell_table(steps) if str(points.get("preset", "")).lower() == "bell" else torch.ones(steps)
    pts = []
    for p in points:
        try:
            pts.append((max(0.0, min(1.0, float(p[0]))), max(0.0, float(p[1]))))
        except (TypeError, ValueError, IndexError):
            pass
    if len(pts) < 2:
        return torch.ones(steps)
    pts.sort(key=lambda p: p[0])
    pts = [(0.0, pts[0][1])] + (pts if pts[0][0] > 0.0 else pts[1:])
    pts = (pts if pts[-1][0] < 1.0 else pts[:-1]) + [(1.0, pts[-1][1])]
    xp = torch.tensor([p[0] for p in pts], dtype=torch.float32)
    yp = torch.tensor([p[1] for p in pts], dtype=torch.float32)
    grid = torch.linspace(0.0, 1.0, steps, dtype=torch.float32)
    idx = torch.searchsorted(xp, grid, right=True).clamp(1, len(pts) - 1)
    blend = ((grid - xp[idx - 1]) / (xp[idx] - xp[idx - 1]).clamp_min(1e-12)).clamp(0.0, 1.0)
    return yp[idx - 1] + (yp[idx] - yp[idx - 1]) * blend


def weighted_mse(pred, target, timesteps, table=None):
    """train.py:2408-2416."""
    per = (pred.float() - target.float()).pow(2).flatten(1).mean(dim=1)
    if table is None:
        return per.mean()
    w = table.to(per.dtype)[timesteps.long().clamp(0, table.shape[0] - 1)]
    return (per * w).mean()


# ----------------------------------------------------------------------------------------
# layer exclusion (train.py:2664-2667)
# ----------------------------------------------------------------------------------------
def is_excluded(name, keywords):
    return any(fnmatch.fnmatch(name, kw if "*" in kw else f"*{kw}*") for kw in keywords)


# ----------------------------------------------------------------------------------------
# Raven / Titan update math (raven.py:89-149, titan.py:237-296) on one tensor
# ----------------------------------------------------------------------------------------
def raven_scalars(lr, betas, eps, weight_decay, debias_strength, step):
    """float64 host scalars exactly as raven.py:101-137 forms them."""
    b1, b2 = betas
    wd_factor = 1.0 - lr * weight_decay if weight_decay != 0 else 1.0
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    if debias_strength < 1.0:
        bc1 = 1.0 - (1.0 - bc1) * debias_strength
        bc2 = 1.0 - (1.0 - bc2) * debias_strength
    return dict(beta1=b1, beta2=b2, one_m_b1=1.0 - b1, one_m_b2=1.0 - b2, eps=eps,
                wd_factor=wd_factor, sqrt_bc2=math.sqrt(bc2), step_size=lr / bc1)


def raven_update_(p, g, m, v, *, lr, betas, eps, weight_decay, debias_strength, step):
    """In-place Raven update of one tensor; fp32 math, stores rounded to p / m dtypes
    (raven.py:122-147).  ``p``: fp32 or bf16, ``g`` same dtype, ``m``/``v``: momentum dtype."""
    s = raven_scalars(lr, betas, eps, weight_decay, debias_strength, step)
    gf = g.float()
    mf = m.float().mul_(s["beta1"]).add_(gf, alpha=s["one_m_b1"])
    vf = v.float().mul_(s["beta2"]).addcmul_(gf, gf, value=s["one_m_b2"])
    pf = p.float()
    if weight_decay != 0:
        pf.mul_(s["wd_factor"])
    denom = vf.sqrt().div_(s["sqrt_bc2"]).add_(eps)
    pf.addcdiv_(mf, denom, value=-s["step_size"])
    p.copy_(pf)
    m.copy_(mf)
    v.copy_(vf)


def clip_grad_norm_ref(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ semantics (train.py:2775-2778) incl. its dtype behaviour:
    with bf16 grads torch returns a bf16 norm and scales by a bf16 coefficient."""
    params = [torch.nn.Parameter(torch.empty_like(g)) for g in grads]
    for p, g in zip(params, grads):
        p.grad = g
    return torch.nn.utils.clip_grad_norm_(params, max_norm)
