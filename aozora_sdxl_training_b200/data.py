"""Cached-latent data path in front of the training step: the reference's on-disk cache format,
its per-sample conditioning rules and its batch schedules, feeding ``SDXLTrainStep`` from pinned host memory.

What stays identical to the reference (checked item for item against its own classes in ``tests/test_data.py``):

* the cache format: ``<dataset>/.precomputed_embeddings_cache_{rf,standard_sdxl}/dataset_index.pt`` (``{"files": [...]}``),
  ``*_lat.pt`` (tensor or ``{"latents": ...}``), ``*_te.pt`` (``{"embeds", "pooled"}``), ``null_embeds.pt``
  (train.py:1992-2043; training_utils/caching/cache.py:75-98);
* the item order: per dataset the index entries sorted by the stable key (cache.py:113-121), repeated ``repeats`` times,
  then ONE ``random.Random(seed).shuffle`` of the whole list (train.py:2009-2024);
* the per-sample stream: ``random.Random(sha256("{seed}:sdxl-sample:{sample}:{item}")[:8])`` drives, in this order, the
  caption-variant choice, the unconditional-dropout draw and the conditioning-scale draw (train.py:2064-2067, 2118-2161;
  cache.py:218-246);
* null-conditioning alignment for chunked captions (train.py:2079-2116), collate (2213-2222), packed sample indices
  (2045-2062, 2245-2254), ``time_ids`` rows (2726-2731) and the bucketed epoch-shuffle batch schedule (461-534, 767-778).

What is B200-first: ``BatchFeeder`` -- the reference loads every item with ``torch.load`` on the training thread
(``NUM_WORKERS=0``); here a background thread builds the next batches into a ring of PINNED host buffers while the GPU
runs the current step, and ``SDXLTrainStep.step`` copies them with non-blocking H2D transfers.
"""
from __future__ import annotations

import hashlib
import math
import queue
import random
import threading
from collections import defaultdict
from pathlib import Path

import torch

CACHE_INDEX_NAME = "dataset_index.pt"                       # cache.py:13
CAPTION_TYPES = ("tags", "nl", "tags_nl", "nl_tags")        # cache.py:10
CAPTION_PRIMARY = "tags_nl"                                 # cache.py:11
CLIP_CHUNK_TOKENS = 77                                      # train.py:145
INDEX_BITS = 32
INDEX_MASK = (1 << INDEX_BITS) - 1


# ------------------------------------------------------------------------------------------------------------
# config helpers (train.py:86-97, 1227-1237; cache.py:205-215)
# ------------------------------------------------------------------------------------------------------------
def caption_weights(config) -> dict:
    w = {"tags": int(getattr(config, "CAPTION_TAGS_PERCENT", 40) or 0), "nl": int(getattr(config, "CAPTION_NL_PERCENT", 10) or 0),
         "tags_nl": int(getattr(config, "CAPTION_TAGS_NL_PERCENT", 25) or 0), "nl_tags": int(getattr(config, "CAPTION_NL_TAGS_PERCENT", 25) or 0)}
    w = {k: max(0, v) for k, v in w.items()}
    if sum(w.values()) <= 0:
        w[CAPTION_PRIMARY] = 100
    return w


def json_captions(config) -> bool:
    return str(getattr(config, "CAPTION_SOURCE_TYPE", "txt") or "txt").strip().lower() == "json"


def conditioning_scale_range(config):
    if not bool(getattr(config, "TEXT_CONDITIONING_SCALE_ENABLED", False)):
        return 1.0, 1.0
    lo = min(max(float(getattr(config, "TEXT_CONDITIONING_SCALE_MIN", 1.0)), 0.0), 1.0)
    hi = min(max(float(getattr(config, "TEXT_CONDITIONING_SCALE_MAX", 1.0)), 0.0), 2.0)
    return (hi, lo) if lo > hi else (lo, hi)


def stable_item_key(item):
    """Sort key that does not depend on filesystem traversal order (cache.py:113-121)."""
    return (str(item.get("relative_path", item.get("image_key", ""))).replace("\\", "/").casefold(),
            int(item.get("bucket_variant_index", 0) or 0), tuple(item.get("target_size", (0, 0))),
            str(item.get("lat_path", item.get("te_path", ""))).replace("\\", "/").casefold())


def _pick_caption_type(rng, weights):
    total = sum(max(0, int(weights.get(k, 0) or 0)) for k in CAPTION_TYPES)
    if total <= 0:
        return CAPTION_PRIMARY
    roll, upto = rng.uniform(0, total), 0
    for key in CAPTION_TYPES:
        upto += max(0, int(weights.get(key, 0) or 0))
        if roll <= upto:
            return key
    return CAPTION_PRIMARY


def caption_variant_path(item, rng, weights, enabled=True):
    """Text-embedding file of the caption variant drawn for this sample (cache.py:238-246).  Consumes ONE ``rng.uniform``
    exactly when the item carries variants and JSON captions are on -- the later dropout / scale draws depend on that."""
    variants = item.get("caption_variants")
    if enabled and isinstance(variants, dict):
        kind = _pick_caption_type(rng, {k: weights.get(k, 0) for k in variants})
        variant = variants.get(kind) or variants.get(CAPTION_PRIMARY) or next(iter(variants.values()))
        if isinstance(variant, dict) and variant.get("te_path"):
            return variant["te_path"]
    return item.get("te_path")


def pack_sample_index(dataset_index: int, sample_index: int) -> int:
    dataset_index, sample_index = int(dataset_index), int(sample_index)
    if dataset_index < 0 or dataset_index > INDEX_MASK:
        raise ValueError(f"Dataset index is too large to pack deterministically: {dataset_index}")
    return (sample_index << INDEX_BITS) | dataset_index


def unpack_sample_index(packed: int):
    packed = int(packed)
    return packed & INDEX_MASK, packed >> INDEX_BITS


# ------------------------------------------------------------------------------------------------------------
# the dataset
# ------------------------------------------------------------------------------------------------------------
class CachedLatentDataset(torch.utils.data.Dataset):
    """Drop-in for ``ImageTextLatentDataset`` (train.py:1992-2163): same constructor argument, ``items`` / ``bucket_keys``
    order, packed-index ``__getitem__`` and returned dictionary."""

    SAMPLE_INDEX_BITS = INDEX_BITS
    SAMPLE_INDEX_MASK = INDEX_MASK
    pack_sample_index = staticmethod(pack_sample_index)
    unpack_sample_index = staticmethod(unpack_sample_index)

    def __init__(self, config):
        self.seed = config.SEED if config.SEED else 42
        self.json_caption_mode = json_captions(config)
        self.caption_weights = caption_weights(config)
        rf = bool(getattr(config, "is_rectified_flow", False))
        self.cache_folder = ".precomputed_embeddings_cache_rf" if rf else ".precomputed_embeddings_cache_standard_sdxl"
        pairs = []
        for ds in config.INSTANCE_DATASETS:
            cache_dir = Path(ds["path"]) / self.cache_folder
            if not (cache_dir / CACHE_INDEX_NAME).exists():
                print(f"WARNING: Index missing at {cache_dir}. Please re-run caching!")
                continue
            index = torch.load(cache_dir / CACHE_INDEX_NAME, map_location="cpu", weights_only=False)
            ordered = sorted(index["files"], key=stable_item_key)
            for _ in range(int(ds.get("repeats", 1))):
                pairs.extend((item, tuple(item["target_size"])) for item in ordered)
        if not pairs:
            raise ValueError("No cached files found.")
        random.Random(self.seed).shuffle(pairs)
        self.items = [p[0] for p in pairs]
        self.bucket_keys = [p[1] for p in pairs]
        self.null_embeds = self.null_pooled = None
        self.cond_scale_min, self.cond_scale_max = conditioning_scale_range(config)
        self.cond_scale_enabled = self.cond_scale_min < 1.0 or self.cond_scale_max > 1.0
        self.dropout_prob = (min(max(float(getattr(config, "UNCONDITIONAL_DROPOUT_CHANCE", 0.0)), 0.0), 1.0)
                             if getattr(config, "UNCONDITIONAL_DROPOUT", False) else 0.0)
        if self.dropout_prob > 0 or self.cond_scale_enabled:
            try:
                null = torch.load(Path(config.INSTANCE_DATASETS[0]["path"]) / self.cache_folder / "null_embeds.pt", map_location="cpu",
                                  weights_only=True)
                self.null_embeds = null["embeds"].squeeze(0) if null["embeds"].dim() == 3 else null["embeds"]
                self.null_pooled = null["pooled"].squeeze(0) if null["pooled"].dim() == 2 else null["pooled"]
            except Exception:                   # the reference silently trains fully conditioned without the null cache
                self.dropout_prob, self.cond_scale_enabled = 0.0, False

    def __len__(self):
        return len(self.items)

    def sample_rng(self, dataset_index, sample_index):
        digest = hashlib.sha256(f"{self.seed}:sdxl-sample:{int(sample_index)}:{int(dataset_index)}".encode("utf-8")).digest()
        return random.Random(int.from_bytes(digest[:8], "little"))

    # -- null conditioning of a different token length (chunked captions), train.py:2079-2116 -----------------
    def _null_of_length(self, n, dtype):
        null = self.null_embeds
        have = null.shape[0]
        if have == n:
            return null.to(dtype=dtype)
        if n < have:
            return null[:n].to(dtype=dtype)
        chunk = CLIP_CHUNK_TOKENS if have >= CLIP_CHUNK_TOKENS else have
        if chunk <= 0 or have % chunk != 0:
            return torch.cat([null, null[-1:].expand(n - have, -1)], dim=0).to(dtype=dtype)
        last = null[-chunk:]
        whole, part = divmod(n - have, chunk)
        parts = [null] + ([last.repeat(whole, 1)] if whole else []) + ([last[:part]] if part else [])
        return torch.cat(parts, dim=0).to(dtype=dtype)

    def _aligned(self, embeds):
        null = self.null_embeds
        if null is None or embeds.shape == null.shape or embeds.dim() != 2 or null.dim() != 2 or embeds.shape[1] != null.shape[1]:
            return embeds, null
        if embeds.shape[0] < null.shape[0]:
            embeds = torch.cat([embeds, self._null_of_length(null.shape[0], embeds.dtype)[embeds.shape[0]:null.shape[0]]], dim=0)
        elif embeds.shape[0] > null.shape[0]:
            null = self._null_of_length(embeds.shape[0], null.dtype)
        return embeds, null

    def __getitem__(self, packed_index):
        try:
            dataset_index, sample_index = unpack_sample_index(packed_index)
            rng = self.sample_rng(dataset_index, sample_index)
            meta = self.items[dataset_index]
            te_path = caption_variant_path(meta, rng, self.caption_weights, enabled=self.json_caption_mode)
            text = torch.load(te_path, map_location="cpu", weights_only=True)
            latents = torch.load(meta["lat_path"], map_location="cpu", weights_only=True)
            if isinstance(latents, dict):
                latents = latents.get("latents")
            if torch.isnan(latents).any() or torch.isinf(latents).any():
                return None
            embeds, pooled = text["embeds"], text["pooled"]
            item = {"latents": latents,
                    "embeds": embeds.squeeze(0) if embeds.dim() == 3 else embeds,
                    "pooled": pooled.squeeze(0) if pooled.dim() == 2 else pooled,
                    "original_sizes": meta["original_size"],
                    "scaled_sizes": meta.get("scaled_size", meta["original_size"]),
                    "target_sizes": meta["target_size"],
                    "crop_coords": meta.get("crop_coords", (0, 0)),
                    "latent_path": te_path,
                    "image_key": meta.get("relative_path", meta["lat_path"])}
            if self.dropout_prob > 0 and rng.random() < self.dropout_prob:
                _, null = self._aligned(item["embeds"])
                item["embeds"], item["pooled"] = null, self.null_pooled
            elif self.cond_scale_enabled:
                scale = rng.uniform(self.cond_scale_min, self.cond_scale_max)
                embeds, null = self._aligned(item["embeds"])
                item["embeds"] = null + (embeds - null) * scale
                item["pooled"] = self.null_pooled + (item["pooled"] - self.null_pooled) * scale
            return item
        except Exception as e:                  # a broken cache file drops the sample, as in the reference
            print(f"[DATASET] Failed to load item {packed_index}: {e}")
            return None


def collate(batch):
    """custom_collate_fn (train.py:2213-2222): failed items dropped, tensors stacked, everything else listed."""
    batch = [b for b in batch if b]
    if not batch:
        return {}
    out = {}
    for k, first in batch[0].items():
        if k != "original_image" and isinstance(first, torch.Tensor):
            out[k] = torch.stack([b[k] for b in batch])
        else:
            out[k] = [b[k] for b in batch]
    return out


def time_ids_rows(batch):
    """``[scaled_h, scaled_w, crop_top, crop_left, target_h, target_w]`` per sample (train.py:2722-2731); the caller turns
    them into a bf16 tensor, which is part of the numerics contract (1365 becomes 1368)."""
    n = len(batch["latents"])
    crops = batch.get("crop_coords", [(0, 0)] * n)
    scaled = batch.get("scaled_sizes", batch["original_sizes"])
    return [[s[1], s[0], c[0], c[1], t[1], t[0]] for s, c, t in zip(scaled, crops, batch["target_sizes"])]


# ------------------------------------------------------------------------------------------------------------
# batch schedules (train.py:461-534, 767-778, 2245-2254)
# ------------------------------------------------------------------------------------------------------------
def bucket_epoch_batches(bucket_keys, batch_size, seed, epoch, shuffle=True):
    """One epoch of single-bucket batches in the reference's order (BucketBatchSampler.__iter__, train.py:478-534): a seeded
    permutation is cut per bucket into chunks of ``batch_size`` (short chunks anywhere), chunks are shuffled inside their
    bucket and buckets are interleaved, always drawing among the buckets with the most chunks left and never the same
    bucket twice in a row when another one is available.  Every draw comes from ONE ``torch.Generator(seed + epoch)``."""
    g = torch.Generator()
    g.manual_seed(seed + epoch)
    order = torch.randperm(len(bucket_keys), generator=g).tolist()
    if batch_size == 1:
        return [[i] for i in order]
    per_bucket = defaultdict(list)
    for idx in order:
        per_bucket[bucket_keys[idx]].append(idx)
    chunks = {}
    for key in sorted(per_bucket):
        members = per_bucket[key]
        cut = [members[i:i + batch_size] for i in range(0, len(members), batch_size)]
        if shuffle and len(cut) > 1:
            cut = [cut[i] for i in torch.randperm(len(cut), generator=g).tolist()]
        chunks[key] = cut
    if not shuffle:
        return [b for key in sorted(chunks) for b in chunks[key]]
    batches, last = [], None
    while chunks:
        cand = [k for k in chunks if k != last] or list(chunks)
        most = max(len(chunks[k]) for k in cand)
        top = [k for k in cand if len(chunks[k]) == most]
        key = top[torch.randint(len(top), (1,), generator=g).item()]
        batches.append(chunks[key].pop(0))
        last = key
        if not chunks[key]:
            del chunks[key]
    return batches


def epoch_shuffle_batch_schedule(bucket_keys, total_steps, batch_size, seed):
    """``build_epoch_shuffle_batch_schedule`` (train.py:767-778): epochs of ``bucket_epoch_batches`` until ``total_steps``."""
    schedule, epoch = [], 0
    while len(schedule) < total_steps:
        for batch in bucket_epoch_batches(bucket_keys, batch_size, seed, epoch):
            schedule.append([int(i) for i in batch])
            if len(schedule) >= total_steps:
                break
        epoch += 1
    return schedule


def pack_sample_schedule(image_schedule, batch_size):
    """``pack_sdxl_sample_schedule`` (train.py:2245-2254): sample number = batch_index * batch_size + position, whatever the
    batch's actual length."""
    batch_size = max(1, int(batch_size or 1))
    return [[pack_sample_index(d, bi * batch_size + li) for li, d in enumerate(batch)] for bi, batch in enumerate(image_schedule)]


def rank_rows(global_len, rank, world):
    """[lo, hi) = the rows of a ``global_len``-row global batch that rank ``rank`` of ``world`` processes.  A short batch (the
    bucket leftovers of train.py:493-496) gives the last ranks fewer, possibly zero, rows."""
    per = math.ceil(global_len / world) if global_len else 0
    return min(global_len, rank * per), min(global_len, (rank + 1) * per)


def rank_slice(global_batch, rank, world):
    """Data parallel: rank r's rows of a global batch (global-batch equivalence)."""
    lo, hi = rank_rows(len(global_batch), rank, world)
    return global_batch[lo:hi]


# ------------------------------------------------------------------------------------------------------------
# pinned-memory prefetch
# ------------------------------------------------------------------------------------------------------------
class BatchFeeder:
    """Iterates a packed schedule and yields collated batches whose tensors live in PINNED host memory, built ``depth``
    batches ahead by a background thread (file reads and conditioning arithmetic overlap the GPU step; the H2D copies in
    ``SDXLTrainStep.step`` are then truly asynchronous).  ``rank`` / ``world`` select this rank's rows of every global batch."""

    _STOP = object()

    def __init__(self, dataset, packed_schedule, *, depth=3, rank=0, world=1, start_step=0, pin=None):
        self.dataset = dataset
        self.schedule = packed_schedule
        self.depth = max(1, int(depth))
        self.rank, self.world = rank, world
        self.start_step = max(0, int(start_step or 0))
        self.pin = torch.cuda.is_available() if pin is None else pin
        self._q = None
        self._thread = None

    def __len__(self):
        return max(0, len(self.schedule) - self.start_step)

    def _build(self, packed_batch):
        mine = rank_slice(packed_batch, self.rank, self.world) if self.world > 1 else packed_batch
        batch = collate([self.dataset[i] for i in mine])
        if batch:
            batch["time_ids"] = time_ids_rows(batch)
            if self.pin:
                for k, v in batch.items():
                    if isinstance(v, torch.Tensor):
                        batch[k] = v.pin_memory()
        if self.world > 1:
            # every rank must know the size of the GLOBAL batch and where its rows sit in it: tickets, noise rows and the loss
            # normaliser are those of the single-process run (SDXLTrainStep reads these two keys); a rank whose slice is empty
            # still gets a (metadata-only) batch, because it has to take part in the step's collectives
            batch = dict(batch) if batch else {}
            batch["global_batch_len"] = len(packed_batch)
            batch["global_row_offset"] = rank_rows(len(packed_batch), self.rank, self.world)[0]
        return batch

    def _worker(self, q):
        try:
            for step in range(self.start_step, len(self.schedule)):
                q.put(self._build(self.schedule[step]))
        except Exception as e:                  # surface loader failures on the consumer side
            q.put(e)
        q.put(self._STOP)

    def __iter__(self):
        q = queue.Queue(maxsize=self.depth)
        self._q = q
        self._thread = threading.Thread(target=self._worker, args=(q,), daemon=True)
        self._thread.start()
        while True:
            item = q.get()
            if item is self._STOP:
                return
            if isinstance(item, Exception):
                raise item
            yield item


# ------------------------------------------------------------------------------------------------------------
# bucket ladder (train.py:894-999): which (width, height) a source image is cached at
# ------------------------------------------------------------------------------------------------------------
SDXL_BUCKETS = [(1024, 1024), (1152, 896), (896, 1152), (1216, 832), (832, 1216), (1344, 768), (768, 1344), (1440, 720), (720, 1440),
                (1536, 640), (640, 1536), (1600, 512), (512, 1600), (896, 896), (768, 768)]
LOW_RES_BUCKETS = [(1152, 512), (512, 1152), (1024, 576), (576, 1024), (960, 640), (640, 960), (896, 704), (704, 896), (768, 768)]
MAX_BUCKET_CHOICES = (896, 1024, 1152, 1536)


def resolve_max_bucket_resolution(value=None):
    """Largest ladder tier not above ``value`` (an edge length; values above 4096 are read as a pixel area)."""
    try:
        n = 1024 if value is None else int(float(value))
    except (TypeError, ValueError):
        return 1024
    if n > 4096:
        n = int(round(math.sqrt(max(1, n))))
    fits = [c for c in MAX_BUCKET_CHOICES if c <= n]
    return fits[-1] if fits else MAX_BUCKET_CHOICES[0]


def bucket_ladder(max_bucket_resolution=None):
    """All cache resolutions for a maximum tier: the 1024 ladder as listed, other tiers by scaling it to multiples of 64;
    sorted by (area, width, height)."""
    top = resolve_max_bucket_resolution(max_bucket_resolution)
    tiers = [top] if top < 1024 else [1024] + [t for t in (1152, 1536) if t <= top]
    base = SDXL_BUCKETS + LOW_RES_BUCKETS
    out = set()
    for tier in tiers:
        if tier == 1024:
            out.update(base)
        else:
            k = tier / 1024
            out.update((max(64, int(round(w * k / 64)) * 64), max(64, int(round(h * k / 64)) * 64)) for w, h in base)
    return sorted(out, key=lambda b: (b[0] * b[1], b[0], b[1]))


def _bucket_cost(bucket, aspect, area):
    """10 x relative aspect error + |log(area ratio)| (train.py:958-963)."""
    w, h = bucket
    ar_err = abs(w / max(h, 1) - aspect) / max(aspect, 0.01)
    return ar_err * 10.0 + (abs(math.log(w * h / area)) if w * h > 0 else 100.0)


def optimal_bucket(orig_w, orig_h, target_area=None, stride=64, should_upscale=False):
    """Best ladder entry for an image; without upscaling, the largest entry that fits inside the image (or, if none fits, the
    best of the smallest-area entries)."""
    aspect = orig_w / max(orig_h, 1)
    top = resolve_max_bucket_resolution(target_area)
    ladder = bucket_ladder(top)
    area = top * top
    best = min(ladder, key=lambda b: _bucket_cost(b, aspect, area))
    if not should_upscale and (best[0] > orig_w or best[1] > orig_h):
        fitting = [b for b in ladder if b[0] <= orig_w and b[1] <= orig_h]
        if fitting:
            return max(fitting, key=lambda b: b[0] * b[1])
        floor = min(b[0] * b[1] for b in ladder)
        return min((b for b in ladder if b[0] * b[1] <= floor * 1.1), key=lambda b: _bucket_cost(b, aspect, area))
    return best


def multi_bucket_resolutions(orig_w, orig_h, target_area=None, should_upscale=False, max_extra=0):
    """Primary bucket plus up to ``max_extra`` next-best alternatives (multi-bucket caching)."""
    primary = optimal_bucket(orig_w, orig_h, target_area, 64, should_upscale)
    if max_extra <= 0:
        return [primary]
    aspect = orig_w / max(orig_h, 1)
    top = resolve_max_bucket_resolution(target_area)
    ranked = sorted(((_bucket_cost(b, aspect, top * top), b) for b in bucket_ladder(top)
                     if b != primary and (should_upscale or (b[0] <= orig_w and b[1] <= orig_h))), key=lambda t: t[0])
    return [primary] + [b for _, b in ranked[:max_extra]]


# ------------------------------------------------------------------------------------------------------------
# timestep-spread schedules (train.py:565-575, 688-889): besides shuffling, steer which image meets which timestep bin so
# that an image does not see the same bin again within its last few visits
# ------------------------------------------------------------------------------------------------------------
def timestep_bin_ids(timesteps, bin_ranges):
    """Bin number of every ticket (first range with start <= t < end; 0 when none matches)."""
    import numpy as np
    ids = np.zeros(len(timesteps), dtype=np.int32)
    for i, t in enumerate(timesteps):
        t = int(t)
        for b, (lo, hi) in enumerate(bin_ranges):
            if lo <= t < hi:
                ids[i] = b
                break
    return ids


def epoch_shuffle_image_schedule(total_images, total_steps, seed):
    """One image per step: seeded permutations (``torch.Generator(seed + epoch)``) laid end to end (train.py:688-700)."""
    import numpy as np
    out = np.empty(total_steps, dtype=np.uint32)
    filled = epoch = 0
    while filled < total_steps:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(total_images, generator=g).numpy().astype(np.uint32, copy=False)
        take = min(total_images, total_steps - filled)
        out[filled:filled + take] = order[:take]
        filled += take
        epoch += 1
    return out


class _BinHistory:
    """Which timestep bins each image met on its last ``depth`` visits, plus the per-epoch 'not used yet' flags and the
    lazily created candidate queues.  ``pick`` is the selection rule shared by the image and the batch schedule: walk the
    queue of this (scope, bin) from where it stopped and take the first image that is still unused this epoch and has not
    met the bin recently; if the queue runs dry, take -- at random among ties -- the unused image of the fallback pool that
    met the bin least often."""

    def __init__(self, total_images, bin_count, visits_per_image):
        import numpy as np
        self.np = np
        self.depth = max(1, min(bin_count, visits_per_image))
        wide = bin_count >= 255
        self.recent = np.full((total_images, self.depth), 65535 if wide else 255, dtype=np.uint16 if wide else np.uint8)
        self.cursor = np.zeros(total_images, dtype=np.uint16)
        self.total_images = total_images

    def new_epoch(self, seed, epoch):
        np = self.np
        self.unused = np.ones(self.total_images, dtype=np.bool_)
        self.queues, self.where = {}, {}
        self.rng = np.random.Generator(np.random.PCG64(seed + 104729 + epoch))

    def pick(self, key, bin_id, make_queue, fallback_pool):
        np = self.np
        q = self.queues.get(key)
        if q is None:
            q = self.queues[key] = make_queue(self.rng)
            self.where[key] = 0
        pos, chosen = self.where[key], None
        while pos < len(q):
            cand = int(q[pos])
            pos += 1
            if self.unused[cand] and not np.any(self.recent[cand] == bin_id):
                chosen = cand
                break
        self.where[key] = pos
        if chosen is None:
            pool = fallback_pool(self.unused)
            if pool.size == 0:
                return None
            seen = np.count_nonzero(self.recent[pool] == bin_id, axis=1)
            ties = pool[seen == seen.min()]
            chosen = int(ties[int(self.rng.integers(0, len(ties)))])
        self.unused[chosen] = False
        slot = int(self.cursor[chosen] % self.depth)
        self.recent[chosen, slot] = bin_id
        self.cursor[chosen] = (self.cursor[chosen] + 1) % self.depth
        return chosen


def spread_image_schedule(total_images, total_steps, seed, bin_ids, bin_count):
    """``build_spread_image_schedule`` (train.py:703-764): batch size 1, one candidate queue per bin over all images."""
    import numpy as np
    if total_images <= 0 or total_steps <= 0:
        return np.empty(0, dtype=np.uint32)
    if bin_count <= 1:
        return epoch_shuffle_image_schedule(total_images, total_steps, seed)
    hist = _BinHistory(total_images, bin_count, math.ceil(total_steps / total_images))
    out = np.empty(total_steps, dtype=np.uint32)
    done = epoch = 0
    while done < total_steps:
        hist.new_epoch(seed, epoch)
        for step in range(done, done + min(total_images, total_steps - done)):
            b = int(bin_ids[step])
            chosen = hist.pick(b, b, lambda rng: rng.permutation(total_images).astype(np.uint32, copy=False), lambda unused: np.flatnonzero(unused))
            if chosen is None:
                break
            out[step] = chosen
        done += min(total_images, total_steps - done)
        epoch += 1
    return out


def spread_batch_schedule(bucket_keys, total_steps, batch_size, seed, timesteps, bin_ranges):
    """``build_spread_batch_schedule`` (train.py:781-881): the bucketed epoch order decides WHICH bucket and how many samples
    each step takes; the members are then re-picked inside that bucket by the bin-history rule, one candidate queue per
    (bucket, bin)."""
    import numpy as np
    total_images = len(bucket_keys)
    if total_images <= 0 or total_steps <= 0:
        return []
    if batch_size == 1:
        ids = timestep_bin_ids(timesteps, bin_ranges)
        return [[int(i)] for i in spread_image_schedule(total_images, total_steps, seed, ids, len(bin_ranges)).tolist()]
    bin_ids = timestep_bin_ids(timesteps, bin_ranges)
    samples = min(len(timesteps), total_steps * batch_size)
    hist = _BinHistory(total_images, max(1, len(bin_ranges)), math.ceil(samples / total_images))
    members = defaultdict(list)
    for i, key in enumerate(bucket_keys):
        members[key].append(i)
    schedule, offset, epoch = [], 0, 0
    while len(schedule) < total_steps and offset < len(bin_ids):
        hist.new_epoch(seed, epoch)
        for base in bucket_epoch_batches(bucket_keys, batch_size, seed, epoch):
            if len(schedule) >= total_steps:
                break
            key = bucket_keys[base[0]]
            pool = members[key]
            batch = []
            for j in range(len(base)):
                if offset + j >= len(bin_ids):
                    break
                b = int(bin_ids[offset + j])

                def make_queue(rng, pool=pool):
                    q = np.array(pool, dtype=np.uint32)
                    rng.shuffle(q)
                    return q
                chosen = hist.pick((key, b), b, make_queue, lambda unused, pool=pool: np.array([i for i in pool if unused[i]], dtype=np.int64))
                if chosen is None:
                    break
                batch.append(chosen)
            if batch:
                schedule.append(batch)
                offset += len(batch)
            if offset >= len(bin_ids):
                break
        epoch += 1
    return schedule


def image_batch_schedule(bucket_keys, total_steps, batch_size, seed, timesteps, bin_ranges, force_spread):
    """``build_image_batch_schedule`` (train.py:884-887)."""
    if not force_spread:
        return epoch_shuffle_batch_schedule(bucket_keys, total_steps, batch_size, seed)
    return spread_batch_schedule(bucket_keys, total_steps, batch_size, seed, timesteps, bin_ranges)
