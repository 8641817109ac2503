"""Host-side pieces of the training step, same names / arguments / results as the reference's ``train.py``.

These are integer / float64 host computations done once or once per micro-step (SURVEY.md rows a1, a2, a6 tables,
a10, a11); they stay in Python/numpy/torch exactly because their random streams (numpy PCG64, ``random.Random``,
``torch.Generator``) must match the reference bit for bit.  Tickets are required to be bit-exact.

  build_timestep_ticket_pool / TimestepSampler   train.py:577-685, 2163-2210
  generate_noise / seeded_torch_generator        train.py:248-263
  CustomCurveLRScheduler                         train.py:325-359
  bell_timestep_loss_curve / timestep_loss_curve_from_config   train.py:2351-2405
  apply_exclusion                                train.py:2664-2667
  logit_normal_allocation                        gui/gui.py:5594-5603, 2309-2318 (TIMESTEP_ALLOCATION.counts recipe)
"""
from __future__ import annotations

import fnmatch
import math
import random

import numpy as np
import torch

_MASK64 = (1 << 64) - 1


# ------------------------------------------------------------------------------------------------------------
# timestep tickets
# ------------------------------------------------------------------------------------------------------------
def _scale_timestep_counts(counts, target_total):
    """Largest-remainder rescale of integer bin counts to ``target_total`` tickets (ties keep bin order)."""
    target_total = max(0, int(target_total))
    c = np.array([max(0, int(v or 0)) for v in counts], dtype=np.int64)
    total = int(c.sum())
    if target_total <= 0 or total <= 0:
        return [0] * len(c)
    raw = [(int(v) / total) * target_total for v in c]          # python float64 arithmetic, as the reference
    floor = [int(v) for v in raw]
    missing = target_total - sum(floor)
    if missing > 0:
        frac = np.array([r - f for r, f in zip(raw, floor)], dtype=np.float64)
        for idx in np.argsort(-frac, kind="stable")[:missing]:
            floor[int(idx)] += 1
    return floor


def _bins_for(allocation, total_tickets, n_timesteps):
    usable = (bool(allocation) and "counts" in allocation and "bin_size" in allocation
              and sum(allocation["counts"]) != 0)
    if usable:
        bin_size = max(1, int(allocation["bin_size"]))
        counts = _scale_timestep_counts(allocation["counts"], total_tickets)
    else:
        bin_size, nbins = 100, 10
        base, extra = divmod(total_tickets, nbins)
        counts = [base + (1 if i < extra else 0) for i in range(nbins)]
    factor = n_timesteps / 1000.0
    kept_counts, kept_ranges = [], []
    for i, cnt in enumerate(counts):
        if cnt <= 0:
            continue
        lo = int(i * bin_size * factor)
        hi = min(n_timesteps, max(lo + 1, int((i + 1) * bin_size * factor)))
        if lo >= n_timesteps:
            break
        kept_counts.append(int(cnt))
        kept_ranges.append((lo, hi))
    return kept_counts, kept_ranges


def _balanced_order(bin_counts, seed):
    live = [(b, c) for b, c in enumerate(bin_counts) if c > 0]
    if not live:
        return []
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    pos, ids, tie = [], [], []
    for b, c in live:
        pos.append((np.arange(c, dtype=np.float64) + rng.random(c)) / c)     # draw order: position jitter ...
        ids.append(np.full(c, b, dtype=np.int32))
        tie.append(rng.random(c))                                            # ... then tie-break jitter, per bin
    pos, ids, tie = np.concatenate(pos), np.concatenate(ids), np.concatenate(tie)
    return ids[np.lexsort((tie, pos))].tolist()


def _stratified(bin_counts, bin_ranges, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    decks = []
    for cnt, (lo, hi) in zip(bin_counts, bin_ranges):
        values = np.arange(lo, hi, dtype=np.int64)
        deck = []
        while len(deck) < cnt:
            deck.extend(rng.permutation(values).tolist()[: cnt - len(deck)])
        decks.append(deck)
    cursor = [0] * len(decks)
    out = []
    for b in _balanced_order(bin_counts, seed):
        out.append(int(decks[b][cursor[b]]))
        cursor[b] += 1
    return out


def build_timestep_ticket_pool(allocation, total_tickets_needed, total_timestep_count=1000, seed=42, stratified=False):
    """Deterministic pool of integer timesteps; returns ``(pool, bin_ranges)`` (train.py:665-685)."""
    total = max(0, int(total_tickets_needed))
    n_t = max(1, int(total_timestep_count))
    seed = int(seed if seed else 42)
    counts, ranges = _bins_for(allocation, total, n_t)
    if stratified:
        pool = _stratified(counts, ranges, seed)
    else:
        rng = np.random.Generator(np.random.PCG64(seed))
        pool = []
        for cnt, (lo, hi) in zip(counts, ranges):
            pool += rng.integers(lo, hi, size=max(1, int(cnt))).tolist()
        random.Random(seed).shuffle(pool)
    if not pool:
        fb = random.Random(seed)
        pool = [fb.randint(0, n_t - 1) for _ in range(total)]
    while len(pool) < total:
        pool += pool[: total - len(pool)]
    return pool[:total], ranges


class TimestepSampler:
    """Pops ``batch_size`` tickets per micro-step from the precomputed pool (train.py:2163-2210).

    ``config`` needs MAX_TRAIN_STEPS, BATCH_SIZE, SEED, is_rectified_flow and optionally TIMESTEP_ALLOCATION /
    TIMESTEP_STRATIFIED_SAMPLING.  Data-parallel use: build it with the GLOBAL batch size and call
    ``sample_rank(local_batch, rank, world)`` -- rank r receives the tickets the single-process run would have
    given to samples ``[r*b, (r+1)*b)`` of the same micro-step."""

    def __init__(self, config, device):
        self.config = config
        self.device = device
        self.total_tickets_needed = config.MAX_TRAIN_STEPS * config.BATCH_SIZE
        self.seed = config.SEED if config.SEED else 42
        self.is_rectified_flow = config.is_rectified_flow
        self.ticket_pool, self.bin_ranges = build_timestep_ticket_pool(
            getattr(config, "TIMESTEP_ALLOCATION", None), self.total_tickets_needed, 1000, self.seed,
            bool(getattr(config, "TIMESTEP_STRATIFIED_SAMPLING", False)))
        self.pool_index = 0

    def set_current_step(self, micro_step):
        self.pool_index = (micro_step * self.config.BATCH_SIZE) % len(self.ticket_pool)

    def state_dict(self):
        return {"pool_index": self.pool_index}

    def load_state_dict(self, state):
        if isinstance(state, dict):
            self.pool_index = int(state.get("pool_index", self.pool_index)) % len(self.ticket_pool)

    def _pop(self, n):
        out = []
        for _ in range(n):
            if self.pool_index >= len(self.ticket_pool):
                self.pool_index = 0
            out.append(self.ticket_pool[self.pool_index])
            self.pool_index += 1
        return out

    def sample(self, batch_size):
        picked = self._pop(batch_size)
        return torch.tensor(picked, dtype=torch.long, device=self.device), picked[0]

    def sample_rank(self, local_batch, rank, world):
        return self.sample_rows(local_batch * world, rank * local_batch, local_batch)

    def sample_rows(self, global_len, row_offset, rows):
        """Data parallel with short / unequal global batches (train.py:493-496 leftover batches, 2733): EVERY rank pops the
        ``global_len`` tickets the single-process run would pop for this micro-step -- so the pool index stays identical on
        all ranks -- and keeps those of its own rows ``[row_offset, row_offset + rows)`` (possibly none)."""
        picked = self._pop(global_len)
        mine = picked[row_offset:row_offset + rows]
        return torch.tensor(mine, dtype=torch.long, device=self.device), (picked[0] if picked else None)

    def update(self, raw_grad_norm):
        pass


def logit_normal_allocation(mu, sigma, total, bin_size=100):
    """GUI recipe for Logit-Normal TIMESTEP_ALLOCATION counts (gui/gui.py:5594-5603, 2309-2318; gui_math.py:30-46)."""
    nb = math.ceil(1000 / bin_size)

    def cdf(z):
        return 0.5 * (1.0 + math.erf(z / math.sqrt(2.0)))

    def logit(p):
        return math.log(p / (1.0 - p))

    w = []
    for i in range(nb):
        lo, hi = i * bin_size, min(1000, (i + 1) * bin_size)
        za = (logit(max(lo / 1000.0, 1e-6)) - mu) / sigma
        zb = (logit(min(hi / 1000.0, 1.0 - 1e-6)) - mu) / sigma
        w.append(max(0.0, cdf(zb) - cdf(za)))
    s = sum(w)
    exact = [x / s * total for x in w]
    base = [int(math.floor(e)) for e in exact]
    order = sorted(range(nb), key=lambda i: exact[i] - base[i], reverse=True)
    for i in order[: total - sum(base)]:
        base[i] += 1
    return {"bin_size": bin_size, "counts": base}


# ------------------------------------------------------------------------------------------------------------
# generators
# ------------------------------------------------------------------------------------------------------------
def generate_noise(latents, generator, device, dtype=None, step=None, seed=None):
    """fp32 normal noise of ``latents.shape``; reseeds with ``(seed + step) % (2**32 - 1)`` when both are given."""
    if step is not None and seed is not None:
        generator.manual_seed((seed + step) % (2 ** 32 - 1))
    return torch.randn(latents.shape, device=device, dtype=torch.float32, generator=generator)


def seeded_torch_generator(device, seed, *parts):
    """64-bit LCG mix of ``seed`` and ``parts`` -> ``torch.Generator`` (rectified-flow jitter stream)."""
    state = int(seed if seed else 42) & _MASK64
    for part in parts:
        state = (state * 6364136223846793005 + int(part) + 1442695040888963407) & _MASK64
    gen = torch.Generator(device=device)
    gen.manual_seed(state % (2 ** 63 - 1))
    return gen


# ------------------------------------------------------------------------------------------------------------
# LR curve and loss-weight table
# ------------------------------------------------------------------------------------------------------------
class CustomCurveLRScheduler:
    """Piecewise-linear LR over the normalised micro-step; writes ``group['lr'] = lr * lr_scale`` (train.py:325-359)."""

    def __init__(self, optimizer, curve_points, total_micro_steps):
        if not curve_points:
            raise ValueError("LR_CUSTOM_CURVE cannot be empty")
        pts = sorted(curve_points, key=lambda p: p[0])
        if pts[0][0] != 0.0:
            pts.insert(0, [0.0, pts[0][1]])
        if pts[-1][0] != 1.0:
            pts.append([1.0, pts[-1][1]])
        self.optimizer = optimizer
        self.curve_points = pts
        self.total_micro_steps = max(total_micro_steps, 1)
        self.current_micro_step = 0
        self._update_lr()

    def _interpolate_lr(self, x):
        x = max(0.0, min(1.0, x))
        for (x1, y1), (x2, y2) in zip(self.curve_points[:-1], self.curve_points[1:]):
            if x1 <= x <= x2:
                if x2 - x1 == 0:
                    return y1
                return y1 + (x - x1) / (x2 - x1) * (y2 - y1)
        return self.curve_points[-1][1]

    def _update_lr(self):
        lr = self._interpolate_lr(self.current_micro_step / max(self.total_micro_steps - 1, 1))
        for group in self.optimizer.param_groups:
            group["lr"] = lr * group.get("lr_scale", 1.0)

    def step(self, micro_step):
        self.current_micro_step = micro_step
        self._update_lr()

    def get_last_lr(self):
        return [group["lr"] for group in self.optimizer.param_groups]


def bell_timestep_loss_curve(total_timestep_count, device=None, dtype=torch.float32):
    n = int(total_timestep_count)
    grid = torch.arange(n, device=device, dtype=dtype)
    y = torch.exp(-2.0 * ((grid - n / 2) / n).pow(2))
    lo = y.min()
    return (y - lo).clamp_min(0.0) * (n / (y - lo).sum().clamp_min(1e-12))


def timestep_loss_curve_from_config(config, total_timestep_count, device=None, dtype=torch.float32):
    """1000-entry per-timestep loss weight table from ``config.TIMESTEP_LOSS_WEIGHT_CURVE`` (train.py:2360-2405)."""
    n = int(total_timestep_count)
    if n <= 0:
        return torch.ones(1, device=device, dtype=dtype)
    spec = getattr(config, "TIMESTEP_LOSS_WEIGHT_CURVE", None)
    if not spec:
        return torch.ones(n, device=device, dtype=dtype)
    if isinstance(spec, dict):
        if str(spec.get("preset", "")).lower() == "bell":
            return bell_timestep_loss_curve(n, device=device, dtype=dtype)
        return torch.ones(n, device=device, dtype=dtype)
    pts = []
    for item in spec:
        try:
            pts.append((max(0.0, min(1.0, float(item[0]))), max(0.0, float(item[1]))))
        except (TypeError, ValueError, IndexError):
            continue
    if len(pts) < 2:
        return torch.ones(n, device=device, dtype=dtype)
    pts.sort(key=lambda p: p[0])
    pts = ([(0.0, pts[0][1])] + pts) if pts[0][0] > 0.0 else ([(0.0, pts[0][1])] + pts[1:])
    pts = (pts + [(1.0, pts[-1][1])]) if pts[-1][0] < 1.0 else (pts[:-1] + [(1.0, pts[-1][1])])
    xs = torch.tensor([p[0] for p in pts], dtype=torch.float32)
    ys = torch.tensor([p[1] for p in pts], dtype=torch.float32)
    grid = torch.linspace(0.0, 1.0, n, dtype=torch.float32)
    hi = torch.searchsorted(xs, grid, right=True).clamp(1, len(pts) - 1)
    lo = hi - 1
    t = ((grid - xs[lo]) / (xs[hi] - xs[lo]).clamp_min(1e-12)).clamp(0.0, 1.0)
    return (ys[lo] + (ys[hi] - ys[lo]) * t).to(device=device, dtype=dtype)


# ------------------------------------------------------------------------------------------------------------
# layer exclusion
# ------------------------------------------------------------------------------------------------------------
def apply_exclusion(model, exclusion_keywords):
    """``requires_grad = not any(fnmatch(name, kw if '*' in kw else f'*{kw}*'))`` over ``named_parameters()``."""
    frozen = 0
    for name, param in model.named_parameters():
        hit = any(fnmatch.fnmatch(name, kw if "*" in kw else f"*{kw}*") for kw in exclusion_keywords)
        param.requires_grad = not hit
        frozen += param.numel() if hit else 0
    return frozen
