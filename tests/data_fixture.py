"""Synthetic on-disk cache in the reference's format (train.py:1992-2043, cache.py:75-121), shared by tests/test_data.py and
tests/golden/make_data_golden.py so the frozen digests and the live run read byte-identical files."""
import hashlib
import os

import torch

BUCKETS = [(1024, 1024), (896, 1152), (1216, 832)]


class DataCfg:
    """The flat config keys the dataset reads."""
    SEED = 42
    is_rectified_flow = False
    CAPTION_SOURCE_TYPE = "json"
    CAPTION_TAGS_PERCENT = 40
    CAPTION_NL_PERCENT = 10
    CAPTION_TAGS_NL_PERCENT = 25
    CAPTION_NL_TAGS_PERCENT = 25
    UNCONDITIONAL_DROPOUT = True
    UNCONDITIONAL_DROPOUT_CHANCE = 0.3
    TEXT_CONDITIONING_SCALE_ENABLED = True
    TEXT_CONDITIONING_SCALE_MIN = 0.8
    TEXT_CONDITIONING_SCALE_MAX = 1.2
    INSTANCE_DATASETS = None


def build_cache(root, rf=False, n_a=11, n_b=6):
    """Two datasets (the second with repeats=2): three buckets, some items with JSON caption variants, one bucket with
    two-chunk (154-token) text embeddings, one latent with a NaN (must be dropped), a null-conditioning file."""
    folder = ".precomputed_embeddings_cache_rf" if rf else ".precomputed_embeddings_cache_standard_sdxl"
    g = torch.Generator().manual_seed(1234)
    datasets = []
    for name, n, repeats in (("setA", n_a, 1), ("setB", n_b, 2)):
        ds_root = os.path.join(root, name)
        cdir = os.path.join(ds_root, folder)
        os.makedirs(cdir, exist_ok=True)
        files = []
        for i in range(n):
            bucket = (i * 7 + len(name)) % 3
            w, h = BUCKETS[bucket]
            stem = f"img_{i:03d}"
            lat = (torch.randn(4, h // 64, w // 64, generator=g) * 0.8).to(torch.bfloat16)     # tiny latents, same layout rules
            if name == "setA" and i == 4:
                lat[0, 0, 0] = float("nan")
            lat_path = os.path.join(cdir, stem + "_lat.pt")
            torch.save({"latents": lat} if i % 2 == 0 else lat, lat_path)
            tokens = 154 if bucket == 1 else 77                # batches are single-bucket: one token length per batch

            def te(tag):
                p = os.path.join(cdir, f"{stem}{tag}_te.pt")
                emb = torch.randn(tokens, 32, generator=g).to(torch.bfloat16)
                torch.save({"embeds": emb[None] if i % 3 == 0 else emb, "pooled": torch.randn(1, 16, generator=g).to(torch.bfloat16)}, p)
                return p
            item = {"lat_path": lat_path, "te_path": te(""), "original_size": (w + 37 * i, h + 11 * i), "scaled_size": (w + 8, h + 4),
                    "target_size": (w, h), "crop_coords": (i % 3, i % 5), "relative_path": f"Sub/{stem}.PNG"}
            if i % 4 == 1:
                item["caption_variants"] = {k: {"te_path": te("_json_" + k)} for k in ("tags", "nl", "tags_nl", "nl_tags")}
            if i % 6 == 2:
                del item["scaled_size"], item["crop_coords"]
            files.append(item)
        files = files[::-1]                                   # index order must not matter (stable sort key)
        torch.save({"files": files}, os.path.join(cdir, "dataset_index.pt"))
        datasets.append({"path": ds_root, "repeats": repeats})
    first = os.path.join(root, "setA", folder)
    torch.save({"embeds": torch.randn(1, 77, 32, generator=g).to(torch.bfloat16), "pooled": torch.randn(1, 16, generator=g).to(torch.bfloat16)},
               os.path.join(first, "null_embeds.pt"))
    return datasets


def _rel(path, root):
    return os.path.relpath(path, root).replace("\\", "/") if os.path.isabs(path) else path


def item_digest(item, root):
    """Content digest of one dataset item / collated batch entry (paths made relative to the cache root)."""
    if item is None:
        return "None"
    h = hashlib.sha256()
    for k in sorted(item):
        v = item[k]
        h.update(k.encode())
        if isinstance(v, torch.Tensor):
            h.update(str((tuple(v.shape), str(v.dtype))).encode())
            h.update(v.contiguous().view(torch.uint8).numpy().tobytes())
        elif isinstance(v, str):
            h.update(_rel(v, root).encode())
        elif isinstance(v, (list, tuple)) and v and isinstance(v[0], str):
            h.update(repr([_rel(x, root) for x in v]).encode())
        else:
            h.update(repr(v).encode())
    return h.hexdigest()
