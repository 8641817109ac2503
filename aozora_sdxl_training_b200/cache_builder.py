"""Latent / text-embedding cache builder: writes the on-disk cache that ``data.CachedLatentDataset``
(and the reference's ``ImageTextLatentDataset``) trains from.

Replaces ``precompute_and_cache_latents`` / ``check_if_caching_needed`` / ``compute_text_embeddings_sdxl`` of the reference
(train.py:1193-1225, 1285-1436, 1597-1989; helpers in training_utils/caching/cache.py) for the SDXL cache: same folder and
file names, same payload dictionaries key for key (``*_te.pt``, ``*_lat.pt``, ``null_embeds.pt``, ``dataset_index.pt`` version 13),
same bucket assignment, resize / centre-crop geometry, caption sidecars (``.txt`` or ``.json`` with the four variants),
fixed-chunk caption tokenisation, file signatures, cache-option record and reuse rules -- a cache built here is accepted
unchanged by the reference's ``check_if_caching_needed`` and vice versa (``tests/test_cache_builder.py`` runs both on the
same image folder with the same encoders and compares every file).

The encoders are the caller's modules, as in the reference: two tokenizers, two CLIP text encoders
(``te(tokens, output_hidden_states=True)`` -> ``.hidden_states[-2]``, ``[0]``) and a VAE (``vae.encode(x).latent_dist.mean``,
``vae.config.{shift_factor, scaling_factor, latent_channels}``).  Not carried over: the Flux BN32 latent normalisation
(``VAE_NORMALIZATION_MODE="flux_bn32"``, a different model family) -- requesting it raises.

What is B200-first here is the shape of the loop, not the arithmetic.  The reference runs one caption per encoder call,
decodes and resizes the images of a batch on the training thread and writes every ``.pt`` synchronously between two GPU
calls.  Here the three stages overlap:

* image decode + Lanczos resize + crop of batch i+1 run in a thread pool while the VAE encodes batch i;
* all captions of a text batch go through each text encoder as ONE ``[captions x chunks, 77]`` call;
* payloads leave through a bounded writer thread (``torch.save`` to a temporary name, then rename: a killed run never
  leaves a half-written file that the reuse check would have to reject).
"""
from __future__ import annotations

import hashlib
import json
import math
import os
import queue
import re
import threading
from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import torch

from . import data

IMAGE_SUFFIXES = (".jpg", ".jpeg", ".png", ".webp", ".bmp")          # cache.py:9
JSON_CAPTION_KEYS = data.CAPTION_TYPES                                 # ("tags", "nl", "tags_nl", "nl_tags"), cache.py:10
PRIMARY_JSON_CAPTION = data.CAPTION_PRIMARY                            # cache.py:11
INDEX_VERSION = 13                                                     # train.py:1987
BUCKET_LAYOUT = "preset_ladder_v3"                                     # train.py:903
_JSON_TAG = re.compile(r"_json_(tags|nl|tags_nl|nl_tags)$")            # cache.py:12
_BUCKET_TAG = re.compile(r"_mb\d+$")
# which recorded options invalidate what (cache.py:14-42)
_LAYOUT_KEYS = ("cache_schema_version", "bucket_layout", "max_bucket_resolution", "should_upscale", "multi_bucket_enabled",
                "multi_bucket_extra_buckets", "caption_source_type")
_TEXT_KEYS = ("cache_schema_version", "text_cache_float_dtype", "caption_source_type", "caption_json_types",
              "caption_chunking_enabled", "caption_embedding_layout")
_LATENT_KEYS = ("cache_schema_version", "vae_cache_float_dtype", "vae_normalization_mode", "vae_shift_factor", "vae_scaling_factor",
                "vae_latent_channels", "vae_path", "vae_source_path", "vae_source_size", "vae_source_mtime_ns")


# ------------------------------------------------------------------------------------------------------------
# configuration -> names, dtypes, the recorded option set
# ------------------------------------------------------------------------------------------------------------
def cache_folder_name(config) -> str:
    return ".precomputed_embeddings_cache_rf" if getattr(config, "is_rectified_flow", False) else ".precomputed_embeddings_cache_standard_sdxl"


def caption_mode(config_or_value=None) -> str:
    """"json" or "txt" (anything else reads as "txt"), from a config object or a plain string (cache.py:205-210)."""
    v = config_or_value
    if v is not None and not isinstance(v, str):
        v = getattr(v, "CAPTION_SOURCE_TYPE", "txt")
    return "json" if str(v or "txt").strip().lower() == "json" else "txt"


_DTYPE_ALIASES = {"fp32": "float32", "float": "float32", "bf16": "bfloat16", "bfp16": "bfloat16", "fp16": "float16", "half": "float16"}


def storage_dtype(config, key):
    """TEXT_CACHE_PRECISION / VAE_CACHE_PRECISION -> torch dtype; unknown names fall back to bfloat16 (train.py:148-163)."""
    name = str(getattr(config, key, "bfloat16") or "bfloat16").strip().lower()
    name = _DTYPE_ALIASES.get(name, name)
    return {"float32": torch.float32, "float16": torch.float16}.get(name, torch.bfloat16)


def _dtype_name(dt) -> str:
    return str(dt).replace("torch.", "")


def chunking_enabled(config) -> bool:
    return bool(getattr(config, "CAPTION_CHUNKING_ENABLED", False))


def null_cache_wanted(config) -> bool:
    """Unconditional dropout or a conditioning-scale range other than [1, 1] needs the empty-prompt embedding (train.py:1227-1243)."""
    if bool(getattr(config, "UNCONDITIONAL_DROPOUT", False)):
        return True
    lo, hi = data.conditioning_scale_range(config)
    return lo < 1.0 or hi > 1.0


def _max_bucket_area(config) -> int:
    edge = data.resolve_max_bucket_resolution(getattr(config, "MAX_BUCKET_RESOLUTION", 1024))
    return edge * edge


def _extra_buckets(config) -> int:
    if not getattr(config, "MULTI_BUCKET_ENABLED", False):
        return 0
    return max(0, int(getattr(config, "MULTI_BUCKET_EXTRA_BUCKETS", 0) or 0))


def _vae_source(config):
    p = getattr(config, "VAE_PATH", None)
    return p if p and Path(p).exists() else getattr(config, "SINGLE_FILE_CHECKPOINT_PATH", None)


def cache_options(config) -> dict:
    """The option record stored in every payload and in the index (train.py:1245-1283): reuse decisions compare subsets of it."""
    src, src_path, src_size, src_mtime = _vae_source(config), "", None, None
    if src:
        try:
            resolved = Path(src).resolve()
            src_path = str(resolved)
            if resolved.exists():
                st = resolved.stat()
                src_size, src_mtime = st.st_size, st.st_mtime_ns
        except OSError:
            src_path = str(src)
    multi = bool(getattr(config, "MULTI_BUCKET_ENABLED", False))
    return {
        "version": INDEX_VERSION,
        "cache_schema_version": 1,
        "bucket_layout": BUCKET_LAYOUT,
        "text_cache_float_dtype": _dtype_name(storage_dtype(config, "TEXT_CACHE_PRECISION")),
        "vae_cache_float_dtype": _dtype_name(storage_dtype(config, "VAE_CACHE_PRECISION")),
        "max_bucket_resolution": data.resolve_max_bucket_resolution(getattr(config, "MAX_BUCKET_RESOLUTION", 1024)),
        "should_upscale": bool(getattr(config, "SHOULD_UPSCALE", False)),
        "caption_embedding_layout": "fixed_total_chunks",
        "caption_source_type": caption_mode(config),
        "caption_json_types": list(JSON_CAPTION_KEYS),
        "caption_chunking_enabled": chunking_enabled(config),
        "multi_bucket_enabled": multi,
        "multi_bucket_extra_buckets": int(getattr(config, "MULTI_BUCKET_EXTRA_BUCKETS", 0) or 0) if multi else 0,
        "vae_normalization_mode": getattr(config, "VAE_NORMALIZATION_MODE", "scalar"),
        "vae_shift_factor": getattr(config, "VAE_SHIFT_FACTOR", None),
        "vae_scaling_factor": getattr(config, "VAE_SCALING_FACTOR", None),
        "vae_latent_channels": getattr(config, "VAE_LATENT_CHANNELS", None),
        "vae_path": str(getattr(config, "VAE_PATH", "") or ""),
        "vae_source_path": src_path,
        "vae_source_size": src_size,
        "vae_source_mtime_ns": src_mtime,
    }


def _same_options(recorded, wanted, keys) -> bool:
    return isinstance(recorded, dict) and isinstance(wanted, dict) and all(recorded.get(k) == wanted.get(k) for k in keys)


# ------------------------------------------------------------------------------------------------------------
# files: discovery, signatures, captions, cache names
# ------------------------------------------------------------------------------------------------------------
def list_images(root):
    """Every image under ``root``, ordered by case-folded relative POSIX path (cache.py:101-110)."""
    root = Path(root)
    found = [p for ext in IMAGE_SUFFIXES for p in root.rglob(f"*{ext}")]
    return sorted(found, key=lambda p: p.relative_to(root).as_posix().casefold())


def stat_signature(path) -> dict:
    path = Path(path)
    if not path.exists():
        return {"exists": False, "path": str(path)}
    st = path.stat()
    return {"exists": True, "path": str(path), "size": st.st_size, "mtime_ns": st.st_mtime_ns}


def caption_sidecar(image_path, mode="txt") -> Path:
    return Path(image_path).with_suffix(".json" if caption_mode(mode) == "json" else ".txt")


def caption_signature_of_file(image_path, mode="txt") -> dict:
    sig = stat_signature(caption_sidecar(image_path, mode))
    sig["mode"] = caption_mode(mode)
    return sig


def read_captions(image_path, mode="txt") -> dict:
    """{"txt": caption} from the ``.txt`` sidecar (file stem with underscores as spaces when missing or empty), or the
    non-empty variants of the ``.json`` sidecar (train.py:1105-1131)."""
    image_path = Path(image_path)
    if caption_mode(mode) == "json":
        side = image_path.with_suffix(".json")
        if not side.exists():
            raise FileNotFoundError(f"JSON caption sidecar not found: {side}")
        with open(side, "r", encoding="utf-8") as f:
            doc = json.load(f)
        if not isinstance(doc, dict):
            raise ValueError(f"JSON caption must be an object: {side}")
        found = {k: doc[k].strip() for k in JSON_CAPTION_KEYS if isinstance(doc.get(k), str) and doc[k].strip()}
        if not found:
            raise ValueError(f"JSON caption {side} must contain at least one non-empty caption key: {', '.join(JSON_CAPTION_KEYS)}")
        return found
    text = image_path.stem.replace("_", " ")
    side = image_path.with_suffix(".txt")
    if side.exists():
        with open(side, "r", encoding="utf-8", errors="ignore") as f:
            body = f.read().strip()
        if body:
            text = body
    return {"txt": text}


def captions_digest(variants: dict) -> str:
    """sha256 of the variants as compact, key-sorted JSON (train.py:1095-1098)."""
    raw = json.dumps({k: variants[k] for k in sorted(variants)}, ensure_ascii=False, sort_keys=True, separators=(",", ":"))
    return hashlib.sha256(raw.encode("utf-8")).hexdigest()


def image_stem(root, image_path) -> str:
    """Cache file stem of an image: its relative path without suffix, separators as underscores (cache.py:164-165)."""
    return str(Path(image_path).relative_to(root).with_suffix("")).replace(os.sep, "_")


def _item_stem(te_path):
    """``<stem>[_mbK]`` of a text cache file, JSON variant tag removed; None for other files (cache.py:168-173)."""
    name = Path(te_path).name
    return _JSON_TAG.sub("", name[:-len("_te.pt")]) if name.endswith("_te.pt") else None


def _base_stem(path):
    """Image stem a cache file belongs to (bucket-variant and JSON tags removed); None for unrelated files."""
    name = Path(path).name
    if name.endswith("_te.pt"):
        return _BUCKET_TAG.sub("", _item_stem(path))
    if name.endswith("_lat.pt"):
        return _BUCKET_TAG.sub("", name[:-len("_lat.pt")])
    return None


def cache_paths(root, cache_dir, entry, caption_keys, json_mode):
    """({caption key: text file}, latent file) of one planned entry (cache.py:337-343)."""
    stem = image_stem(root, entry["ip"]) + entry.get("cache_suffix", "")
    tag = (lambda k: f"_json_{k}") if json_mode else (lambda k: "")
    return {k: Path(cache_dir) / f"{stem}{tag(k)}_te.pt" for k in caption_keys}, Path(cache_dir) / f"{stem}_lat.pt"


def _te_paths_of(item):
    variants = item.get("caption_variants")
    if isinstance(variants, dict):
        return [v["te_path"] for v in variants.values() if isinstance(v, dict) and v.get("te_path")]
    return [item["te_path"]] if item.get("te_path") else []


def load_index(cache_dir):
    return torch.load(Path(cache_dir) / data.CACHE_INDEX_NAME, map_location="cpu", weights_only=False)


def save_index(cache_dir, payload) -> Path:
    cache_dir = Path(cache_dir)
    cache_dir.mkdir(parents=True, exist_ok=True)
    final = cache_dir / data.CACHE_INDEX_NAME
    tmp = final.with_suffix(final.suffix + ".tmp")
    torch.save(payload, tmp)
    tmp.replace(final)
    return final


# ------------------------------------------------------------------------------------------------------------
# planning: image -> bucket variants with resize / crop geometry
# ------------------------------------------------------------------------------------------------------------
def _geometry(orig_w, orig_h, target_w, target_h):
    """Cover-scale the image over the bucket, centre crop: (scaled size, (top, left)) (train.py:1001-1016, 1060-1066)."""
    k = max(target_w / max(orig_w, 1), target_h / max(orig_h, 1))
    sw, sh = int(round(orig_w * k)), int(round(orig_h * k))
    return (sw, sh), (max(0, (sh - target_h) // 2), max(0, (sw - target_w) // 2))


def plan_image(path, max_area, should_upscale, mode):
    """Probe one image (verify, load, size), read its captions and assign the primary bucket; None if unreadable."""
    from PIL import Image
    try:
        with Image.open(path) as im:
            im.verify()
        with Image.open(path) as im:
            im.load()
            w, h = im.size
        if w <= 0 or h <= 0:
            return None
        tw, th = data.optimal_bucket(w, h, max_area, 64, should_upscale)
        scaled, crop = _geometry(w, h, tw, th)
        variants = read_captions(path, mode)
        main = variants.get("txt") or variants.get(PRIMARY_JSON_CAPTION) or next(iter(variants.values()))
        return {"ip": path, "caption": main, "caption_variants": variants, "caption_signature": captions_digest(variants),
                "target_resolution": (tw, th), "original_size": (w, h), "scaled_size": scaled, "crop_coords": crop,
                "original_area": w * h, "target_area": tw * th, "was_upscaled": should_upscale and (w * h) < max_area}
    except Exception as e:                                   # unreadable image or caption: skipped, like the reference
        print(f"\n[CORRUPT IMAGE OR READ ERROR] Skipping {path}, Reason: {e}")
        return None


def bucket_variant(base, target_w, target_h, index=0):
    w, h = base["original_size"]
    scaled, crop = _geometry(w, h, target_w, target_h)
    out = dict(base)
    out.update(target_resolution=(target_w, target_h), scaled_size=scaled, crop_coords=crop, bucket_variant_index=index,
               cache_suffix="" if index == 0 else f"_mb{index}")
    return out


def fit_image(img, target_w, target_h):
    """RGB, Lanczos cover-resize, centre crop to exactly the bucket (train.py:240-246, 1018-1039)."""
    from PIL import Image
    if img.mode == "P" and "transparency" in img.info:
        img = img.convert("RGBA")
    img = img.convert("RGB")
    w, h = img.size
    k = max(target_w / max(w, 1), target_h / max(h, 1))
    nw, nh = max(int(round(w * k)), target_w), max(int(round(h * k)), target_h)
    img = img.resize((nw, nh), Image.Resampling.LANCZOS)
    left, top = (nw - target_w) // 2, (nh - target_h) // 2
    return img.crop((left, top, left + target_w, top + target_h))


def _to_model_input(img):
    """PIL RGB -> float32 [3, h, w] in [-1, 1] (ToTensor + Normalize(0.5, 0.5), train.py:1676)."""
    import numpy as np
    a = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1)
    return (a.to(torch.float32).div(255) - 0.5) / 0.5


# ------------------------------------------------------------------------------------------------------------
# text: fixed-chunk tokenisation and the two-encoder embedding
# ------------------------------------------------------------------------------------------------------------
def tokenizer_window(tokenizer) -> int:
    return int(getattr(tokenizer, "model_max_length", 77) or 77)


def caption_ids(tokenizer, caption):
    ids = tokenizer(caption, add_special_tokens=False, truncation=False).input_ids
    return ids[0] if ids and isinstance(ids[0], list) else ids


def chunk_count(caption, tokenizer) -> int:
    return max(1, math.ceil(len(caption_ids(tokenizer, caption)) / max(1, tokenizer_window(tokenizer) - 2)))


def chunked_tokens(tokenizer, caption, total_chunks):
    """[total_chunks, window] token rows: BOS + up to window-2 caption tokens + EOS + padding per row (train.py:1176-1190)."""
    window = tokenizer_window(tokenizer)
    body = max(1, window - 2)
    bos, eos = tokenizer.bos_token_id, tokenizer.eos_token_id
    pad = tokenizer.pad_token_id if tokenizer.pad_token_id is not None else eos
    ids = caption_ids(tokenizer, caption)
    rows = []
    for c in range(max(1, int(total_chunks or 1))):
        row = [bos] + ids[c * body:(c + 1) * body] + [eos]
        rows.append((row + [pad] * (window - len(row)))[:window])
    return torch.tensor(rows, dtype=torch.long)


def embed_captions(captions, tok1, tok2, te1, te2, device, chunked=False, total_chunks=None):
    """(embeds [n, chunks*77, d1+d2], pooled [n, d2]) for a list of captions (train.py:1193-1225).  All captions go through each
    encoder in ONE call; rows are regrouped per caption afterwards (penultimate hidden state of both encoders concatenated along
    features, pooled output of the second encoder's FIRST chunk)."""
    captions = list(captions)
    with torch.no_grad():
        if chunked:
            if total_chunks is None:
                total_chunks = max(1, *(max(chunk_count(c, tok1), chunk_count(c, tok2)) for c in captions))
            total_chunks = max(1, int(total_chunks))
            ids1 = torch.cat([chunked_tokens(tok1, c, total_chunks) for c in captions]).to(device)
            ids2 = torch.cat([chunked_tokens(tok2, c, total_chunks) for c in captions]).to(device)
        else:
            total_chunks = 1
            enc = lambda tok: tok(captions, padding="max_length", max_length=tok.model_max_length, truncation=True,  # noqa: E731
                                  return_tensors="pt").input_ids.to(device)
            ids1, ids2 = enc(tok1), enc(tok2)
        out1 = te1(ids1, output_hidden_states=True)
        out2 = te2(ids2, output_hidden_states=True)
        h1, h2 = out1.hidden_states[-2], out2.hidden_states[-2]
        n = len(captions)
        h1 = h1.reshape(n, -1, h1.shape[-1])
        h2 = h2.reshape(n, -1, h2.shape[-1])
        pooled = out2[0].reshape(n, total_chunks, -1)[:, 0]
        return torch.cat([h1, h2], dim=-1), pooled


# ------------------------------------------------------------------------------------------------------------
# reuse rules
# ------------------------------------------------------------------------------------------------------------
def _same_entry(payload, root, entry) -> bool:
    """Does a payload describe this planned entry (cache.py:346-356)?"""
    if not isinstance(payload, dict):
        return False
    return (payload.get("relative_path") == str(entry["ip"].relative_to(root))
            and tuple(payload.get("original_size", ())) == tuple(entry["original_size"])
            and tuple(payload.get("scaled_size", payload.get("original_size", ()))) == tuple(entry.get("scaled_size", entry["original_size"]))
            and tuple(payload.get("target_size", ())) == tuple(entry["target_resolution"])
            and tuple(payload.get("crop_coords", (0, 0))) == tuple(entry.get("crop_coords", (0, 0)))
            and int(payload.get("bucket_variant_index", 0) or 0) == int(entry.get("bucket_variant_index", 0) or 0))


def text_file_reusable(path, root, entry, key, caption, dtype, options) -> bool:
    try:
        p = torch.load(path, map_location="cpu", weights_only=True)
        e, q = p.get("embeds"), p.get("pooled")
        return (e is not None and q is not None and e.dtype == dtype and q.dtype == dtype and p.get("caption_type") == key
                and p.get("caption") == caption and p.get("caption_signature") == entry.get("caption_signature")
                and _same_entry(p, root, entry) and _same_options(p.get("cache_options"), options, _TEXT_KEYS))
    except Exception:
        return False


def latent_file_reusable(path, root, entry, dtype, options) -> bool:
    try:
        p = torch.load(path, map_location="cpu", weights_only=True)
        if not isinstance(p, dict) or not _same_entry(p, root, entry) or not _same_options(p.get("cache_options"), options, _LATENT_KEYS):
            return False
        lat = p.get("latents")
        return lat is not None and lat.dtype == dtype and not torch.isnan(lat).any() and not torch.isinf(lat).any()
    except Exception:
        return False


def _unlink(path):
    try:
        Path(path).unlink()
    except OSError as e:
        print(f"WARNING: Could not remove stale cache file {path}: {e}")


# ------------------------------------------------------------------------------------------------------------
# is the cache current?  (train.py:1285-1436)
# ------------------------------------------------------------------------------------------------------------
def cache_needs_build(config, include_null_cache=True) -> bool:
    folder, options, mode = cache_folder_name(config), cache_options(config), caption_mode(config)
    json_mode = mode == "json"
    stale = False
    if include_null_cache and null_cache_wanted(config):
        if any(ds.get("path") and not (Path(ds["path"]) / folder / "null_embeds.pt").exists() for ds in config.INSTANCE_DATASETS):
            stale = True
    for ds in config.INSTANCE_DATASETS:
        root = Path(ds["path"])
        if not root.exists():
            continue
        cache_dir = root / folder
        images = list_images(root)
        if not images:                                        # an emptied dataset with cache files left behind must be cleaned
            if cache_dir.exists() and list(cache_dir.glob("*_te.pt")):
                stale = True
            elif (cache_dir / data.CACHE_INDEX_NAME).exists():
                try:
                    stale = stale or bool(load_index(cache_dir).get("files"))
                except Exception:
                    stale = True
            continue
        stems = {image_stem(root, p) for p in images}
        if not cache_dir.exists() or not (cache_dir / data.CACHE_INDEX_NAME).exists():
            stale = True
            continue
        try:
            index = load_index(cache_dir)
            files = index.get("files", [])
            indexed = {_base_stem(p) for it in files for p in _te_paths_of(it)} - {None}
            if (not _same_options(index.get("cache_options"), options, _LAYOUT_KEYS) or any("scaled_size" not in it for it in files)
                    or len(files) < len(images) or indexed != stems):
                stale = True
            for it in files:
                te_files, lat_file = _te_paths_of(it), it.get("lat_path")
                if not te_files or not lat_file or not Path(lat_file).exists() or any(not Path(p).exists() for p in te_files):
                    stale = True
                    break
                try:
                    if any(not _same_options(torch.load(p, map_location="cpu", weights_only=True).get("cache_options"), options, _TEXT_KEYS)
                           for p in te_files):
                        stale = True
                        break
                    lat_payload = torch.load(lat_file, map_location="cpu", weights_only=True)
                    if not isinstance(lat_payload, dict) or not _same_options(lat_payload.get("cache_options"), options, _LATENT_KEYS):
                        stale = True
                        break
                except Exception:
                    stale = True
                    break
                rel = it.get("relative_path")
                if rel:
                    try:
                        img = root / rel
                        isig, csig = it.get("image_file_signature"), it.get("caption_file_signature")
                        if isig and csig:
                            if isig != stat_signature(img) or csig != caption_signature_of_file(img, mode):
                                stale = True
                                break
                        elif captions_digest(read_captions(img, mode)) != it.get("caption_signature"):
                            stale = True
                            break
                    except Exception:
                        stale = True
                        break
        except Exception:
            stale = True
        on_disk = list(cache_dir.glob("*_te.pt"))
        if ({_base_stem(f) for f in on_disk} - {None}) != stems:
            stale = True
        try:
            from PIL import Image
            area, extra, upscale = _max_bucket_area(config), _extra_buckets(config), getattr(config, "SHOULD_UPSCALE", False)
            want = 0
            for p in images:
                n_caps = len(read_captions(p, mode)) if json_mode else 1
                with Image.open(p) as im:
                    want += n_caps * len(data.multi_bucket_resolutions(im.width, im.height, area, upscale, extra))
        except Exception:
            stale, want = True, len(images)
        if len(on_disk) < want:
            stale = True
        else:
            for f in on_disk[:10]:                            # spot check of the files themselves
                try:
                    if not _same_options(torch.load(f, map_location="cpu", weights_only=True).get("cache_options"), options, _TEXT_KEYS):
                        stale = True
                        break
                except Exception:
                    stale = True
                    break
    return stale


# ------------------------------------------------------------------------------------------------------------
# the builder
# ------------------------------------------------------------------------------------------------------------
class _Writer:
    """Bounded background ``torch.save``: temporary name first, then rename."""

    def __init__(self, depth=64):
        self.q = queue.Queue(maxsize=depth)
        self.error = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        while True:
            job = self.q.get()
            try:
                if job is None:
                    return
                payload, path = job
                tmp = Path(str(path) + ".tmp")
                torch.save(payload, tmp)
                tmp.replace(path)
            except Exception as e:                            # surfaced by flush() / close()
                self.error = e
            finally:
                self.q.task_done()

    def put(self, payload, path):
        if self.error is not None:
            raise self.error
        self.q.put((payload, path))

    def flush(self):
        """Block until every queued payload is on disk."""
        self.q.join()
        if self.error is not None:
            raise self.error

    def close(self):
        self.q.put(None)
        self.t.join()
        if self.error is not None:
            raise self.error


def build_cache(config, tok1, tok2, te1, te2, vae, device, *, decode_workers=8, text_batch=None, progress=None):
    """Bring the cache of every dataset in ``config.INSTANCE_DATASETS`` up to date (files that are still valid are kept).

    Same arguments as the reference's ``precompute_and_cache_latents``; ``progress(kind, done, total)`` is an optional callback."""
    if not cache_needs_build(config):
        return
    mode = caption_mode(config)
    json_mode = mode == "json"
    norm_mode = str(getattr(config, "VAE_NORMALIZATION_MODE", "scalar")).lower()
    if norm_mode != "scalar":
        raise RuntimeError(f"VAE_NORMALIZATION_MODE={norm_mode!r}: only the scalar (shift, scale) latent normalisation is built here")
    folder, options = cache_folder_name(config), cache_options(config)
    text_dtype, lat_dtype = storage_dtype(config, "TEXT_CACHE_PRECISION"), storage_dtype(config, "VAE_CACHE_PRECISION")
    vae.to(device, dtype=torch.float32)
    for switch in ("enable_tiling", "enable_slicing"):
        if hasattr(vae, switch):
            getattr(vae, switch)()
    shift, scale = getattr(vae.config, "shift_factor", None), vae.config.scaling_factor
    te1.to(device)
    te2.to(device)
    chunked = chunking_enabled(config)
    total_chunks = _dataset_chunk_count(config, tok1, tok2, mode) if chunked else 1

    null = None
    if null_cache_wanted(config):
        e, p = embed_captions([""], tok1, tok2, te1, te2, device, chunked, total_chunks)
        null = {"embeds": e.to(dtype=text_dtype).cpu(), "pooled": p.to(dtype=text_dtype).cpu()}
        if not cache_needs_build(config, include_null_cache=False):          # everything else is current: only the null file is missing
            for ds in config.INSTANCE_DATASETS:
                cdir = Path(ds["path"]) / folder
                if (cdir / data.CACHE_INDEX_NAME).exists() and not (cdir / "null_embeds.pt").exists():
                    torch.save(null, cdir / "null_embeds.pt")
            return

    caption_keys = JSON_CAPTION_KEYS if json_mode else ("txt",)
    batch = max(1, int(getattr(config, "CACHING_BATCH_SIZE", 1) or 1))
    pool = ThreadPoolExecutor(max_workers=max(1, decode_workers))
    writer = _Writer()
    try:
        for ds in config.INSTANCE_DATASETS:
            root = Path(ds["path"])
            cdir = root / folder
            cdir.mkdir(exist_ok=True)
            images = list_images(root)
            stems = {image_stem(root, p) for p in images}
            rebuild = bool(getattr(config, "REBUILD_CACHE", False))
            # files of images that no longer exist (or everything, on a forced rebuild)
            doomed = (list(cdir.glob("*_te*.pt")) + list(cdir.glob("*_lat.pt"))) if rebuild else \
                [f for f in cdir.glob("*.pt") if f.name not in ("null_embeds.pt", data.CACHE_INDEX_NAME) and _base_stem(f) not in stems]
            for f in doomed:
                if f.exists():
                    _unlink(f)
            if null is not None:
                torch.save(null, cdir / "null_embeds.pt")

            entries = []
            if images:
                area, extra, upscale = _max_bucket_area(config), _extra_buckets(config), getattr(config, "SHOULD_UPSCALE", False)
                for base in pool.map(lambda p: plan_image(p, area, upscale, mode), images):
                    if base is None:
                        continue
                    w, h = base["original_size"]
                    for k, (tw, th) in enumerate(data.multi_bucket_resolutions(w, h, area, upscale, extra)):
                        entries.append(bucket_variant(base, tw, th, k))
            text_jobs, lat_jobs, expected = [], [], set()
            for e in entries:
                variants = e.get("caption_variants") or {"txt": e["caption"]}
                keys = tuple(k for k in caption_keys if k in variants)
                te_files, lat_file = cache_paths(root, cdir, e, keys, json_mode)
                expected.add(lat_file.resolve())
                for k in keys:
                    expected.add(te_files[k].resolve())
                    if not (te_files[k].exists() and text_file_reusable(te_files[k], root, e, k, variants[k], text_dtype, options)):
                        text_jobs.append((e, k, variants[k], te_files[k]))
                if not (lat_file.exists() and latent_file_reusable(lat_file, root, e, lat_dtype, options)):
                    lat_jobs.append((e, lat_file))
            for f in cdir.glob("*.pt"):                       # variants of current images that the plan no longer contains
                if f.name not in ("null_embeds.pt", data.CACHE_INDEX_NAME) and _base_stem(f) in stems and f.resolve() not in expected:
                    _unlink(f)

            def common(e):
                w, h = e["target_resolution"]
                return {"relative_path": str(e["ip"].relative_to(root)), "image_file_signature": stat_signature(e["ip"]),
                        "caption_file_signature": caption_signature_of_file(e["ip"], mode), "original_size": e["original_size"],
                        "scaled_size": e.get("scaled_size", e["original_size"]), "target_size": (w, h),
                        "crop_coords": e.get("crop_coords", (0, 0)), "bucket_variant_index": e.get("bucket_variant_index", 0),
                        "caption_signature": e.get("caption_signature"), "cache_options": options, "vae_normalization_mode": norm_mode,
                        "vae_shift": shift, "vae_scale": scale, "flux_bn_eps": None}

            # ---- text: one call per encoder for a whole batch of captions -------------------------------------------------
            tb = text_batch or batch * len(caption_keys)
            for i in range(0, len(text_jobs), tb):
                jobs = text_jobs[i:i + tb]
                embeds, pooled = embed_captions([c for _, _, c, _ in jobs], tok1, tok2, te1, te2, device, chunked, total_chunks)
                embeds, pooled = embeds.to(dtype=text_dtype).cpu(), pooled.to(dtype=text_dtype).cpu()
                for j, (e, key, caption, path) in enumerate(jobs):
                    rec = common(e)
                    rec.update(original_stem=e["ip"].stem, caption_type=key, caption=caption, embeds=embeds[j].clone(), pooled=pooled[j].clone())
                    writer.put(rec, path)
                if progress:
                    progress("text", min(i + tb, len(text_jobs)), len(text_jobs))

            # ---- latents: decode / resize of the next batch overlaps the VAE call of this one ---------------------------
            by_res = defaultdict(list)
            for e, path in lat_jobs:
                by_res[e["target_resolution"]].append((e, path))
            batches = [(res, group[i:i + batch]) for res, group in by_res.items() for i in range(0, len(group), batch)]

            def load_one(job):
                from PIL import Image
                e, _ = job
                w, h = e["target_resolution"]
                try:
                    with Image.open(e["ip"]) as im:
                        return _to_model_input(fit_image(im, w, h))
                except Exception as err:
                    print(f"[SKIP] {e['ip'].name}: {err}")
                    return None

            pending = [pool.submit(lambda b=b: [load_one(j) for j in b[1]]) for b in batches[:2]]
            done = 0
            for bi, (res, jobs) in enumerate(batches):
                tensors = pending.pop(0).result()
                if bi + 2 < len(batches):
                    pending.append(pool.submit(lambda b=batches[bi + 2]: [load_one(j) for j in b[1]]))
                good = [(t, j) for t, j in zip(tensors, jobs) if t is not None]
                for t, (e, path) in zip(tensors, jobs):
                    if t is None and Path(path).exists():
                        _unlink(path)
                done += len(jobs)
                if not good:
                    continue
                with torch.no_grad():
                    lat = vae.encode(torch.stack([t for t, _ in good]).to(device, dtype=torch.float32)).latent_dist.mean
                    lat = (lat - shift) * scale if shift is not None else lat * scale
                lat = lat.to(dtype=lat_dtype).cpu()
                for j, (_, (e, path)) in enumerate(good):
                    rec = common(e)
                    rec["latents"] = lat[j].clone()
                    writer.put(rec, path)
                if progress:
                    progress("latents", done, len(lat_jobs))
            writer.flush()                                    # the index is built from the files on disk

            # ---- index: from the files now on disk -----------------------------------------------------------------------
            files = []
            grouped = defaultdict(dict)
            for f in cdir.glob("*_te.pt"):
                stem = _item_stem(f)
                if stem is None:
                    continue
                if json_mode:
                    try:
                        kind = torch.load(f, map_location="cpu", weights_only=True).get("caption_type")
                    except Exception as err:
                        print(f"WARNING: Could not inspect cached text file {f}: {err}")
                        continue
                    if kind in JSON_CAPTION_KEYS:
                        grouped[stem][kind] = f
                else:
                    grouped[stem][None] = f
            for stem, kinds in grouped.items():
                first = kinds.get(PRIMARY_JSON_CAPTION) or next(iter(kinds.values())) if json_mode else kinds[None]
                try:
                    if _base_stem(first) not in stems:
                        _remove_pair(first)
                        continue
                    lat_file = cdir / f"{stem}_lat.pt"
                    if not lat_file.exists():
                        print(f"WARNING: Skipping cached text file with missing latent: {first}")
                        continue
                    te = torch.load(first, map_location="cpu", weights_only=True)
                    rel = te.get("relative_path")
                    if rel and not (root / rel).exists():
                        _remove_pair(first)
                        continue
                    item = {"te_path": str(first), "lat_path": str(lat_file), "relative_path": rel,
                            "image_file_signature": te.get("image_file_signature"), "caption_file_signature": te.get("caption_file_signature"),
                            "target_size": te.get("target_size"), "original_size": te.get("original_size"),
                            "scaled_size": te.get("scaled_size", te.get("original_size")), "crop_coords": te.get("crop_coords", (0, 0)),
                            "bucket_variant_index": te.get("bucket_variant_index", 0), "caption_signature": te.get("caption_signature")}
                    if json_mode:
                        item["caption_variants"] = {k: {"te_path": str(kinds[k])} for k in JSON_CAPTION_KEYS if k in kinds}
                    files.append(item)
                except Exception as err:
                    print(f"WARNING: Could not index {first}: {err}")
            save_index(cdir, {"version": INDEX_VERSION, "cache_options": options, "files": files})
    finally:
        pool.shutdown(wait=True)
        writer.close()
        te1.cpu()
        te2.cpu()


def _remove_pair(te_path):
    te_path = Path(te_path)
    stem = _item_stem(te_path)
    for p in (te_path, te_path.with_name(f"{stem}_lat.pt")):
        if p.exists():
            _unlink(p)


def _dataset_chunk_count(config, tok1, tok2, mode) -> int:
    """Largest chunk count over every caption of every dataset: all text embeddings share one token length (train.py:1152-1173)."""
    most = 1
    for ds in config.INSTANCE_DATASETS:
        root = Path(ds["path"])
        if not root.exists():
            continue
        for img in (p for ext in IMAGE_SUFFIXES for p in root.rglob(f"*{ext}")):
            try:
                variants = read_captions(img, mode)
            except Exception as e:
                print(f"[CAPTION READ ERROR] Skipping caption chunk scan for {img}, Reason: {e}")
                continue
            for caption in variants.values():
                most = max(most, chunk_count(caption, tok1), chunk_count(caption, tok2))
    return most
