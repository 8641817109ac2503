"""Freeze what the REFERENCE's own caching pass (train.py:1597-1989 ``precompute_and_cache_latents``) writes for the image folders
and stand-in models of tests/test_cache_builder.py into tests/golden/cache_golden.json (container only: needs /root/reference).

Per configuration and per cache file: a digest of every tensor (dtype, shape, bytes) and of every value that does not depend on the
machine (absolute paths and file timestamps are reduced to what is stable: relative names, sizes).
    python tests/golden/make_cache_golden.py"""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import torch  # noqa: E402


def _stable(v, root):
    if isinstance(v, torch.Tensor):
        return ["tensor", str(v.dtype), list(v.shape), hashlib.sha256(v.contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()]
    if isinstance(v, dict):
        out = {}
        for k in sorted(v):
            if k in ("mtime_ns", "vae_source_mtime_ns"):
                continue
            out[k] = _stable(v[k], root)
        return out
    if isinstance(v, (list, tuple)):
        return [_stable(x, root) for x in v]
    if isinstance(v, str) and root in v:
        return os.path.relpath(v, root).replace(os.sep, "/")
    return v


def digest_cache(cache_dir, root, stable_key):
    out = {}
    for name in sorted(os.listdir(cache_dir)):
        if not name.endswith(".pt"):
            continue
        payload = torch.load(os.path.join(cache_dir, name), map_location="cpu", weights_only=False)
        if name == "dataset_index.pt":
            payload = dict(payload, files=sorted(payload["files"], key=stable_key))
        raw = json.dumps(_stable(payload, root), sort_keys=True, ensure_ascii=False)
        out[name] = hashlib.sha256(raw.encode("utf-8")).hexdigest()
    return out


def main():
    import test_cache_builder as T
    from aozora_sdxl_training_b200 import cache_builder as cb, data
    from oracle import ref_shim
    tr = ref_shim.import_reference_train()
    gold = {}
    for case, opts in T.CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            root = os.path.join(tmp, "ds")
            T.make_folder(root, json_mode=opts.get("CAPTION_SOURCE_TYPE") == "json")
            cfg = T.cfg_for(root, **opts)
            tr.precompute_and_cache_latents(cfg, *T.models(0.1 if case == "txt_plain_shift" else None), "cpu")
            gold[case] = digest_cache(os.path.join(root, cb.cache_folder_name(cfg)), root, data.stable_item_key)
    with open(os.path.join(HERE, "cache_golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in gold.items()})


if __name__ == "__main__":
    main()
