"""Caching pass, wall clock on the host: the reference's precompute_and_cache_latents against cache_builder.build_cache on the
same synthetic image folder with the same stand-in models (tests/test_cache_builder.py), CPU only.  What differs is the loop:
batched text-encoder calls, decode / resize overlapped with the VAE call, background writes.
    python tools/cache_bench.py [n_images]        (needs /root/reference; container only)"""
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402

import test_cache_builder as T  # noqa: E402
from aozora_sdxl_training_b200 import cache_builder as cb  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    tmp = tempfile.mkdtemp(prefix="cache_bench_")
    root = os.path.join(tmp, "ds")
    os.makedirs(root)
    rng = np.random.default_rng(0)
    sizes = [(1400, 1100), (1100, 1500), (1800, 900), (1200, 1200)]
    for i in range(n):
        w, h = sizes[i % len(sizes)]
        arr = rng.integers(0, 256, size=(h // 16, w // 16, 3), dtype=np.uint8)
        Image.fromarray(arr, "RGB").resize((w, h), Image.Resampling.BILINEAR).save(os.path.join(root, f"img_{i:04d}.jpg"), quality=90)
        open(os.path.join(root, f"img_{i:04d}.txt"), "w").write(" ".join(f"tag{(i * 7 + k) % 97}" for k in range(20 + i % 90)))
    cfg = T.cfg_for(root, CACHING_BATCH_SIZE=8, MULTI_BUCKET_ENABLED=False)
    cdir = os.path.join(root, cb.cache_folder_name(cfg))
    rows = []
    if ref_shim.reference_available():
        tr = ref_shim.import_reference_train()
        t0 = time.perf_counter()
        tr.precompute_and_cache_latents(cfg, *T.models(), "cpu")
        rows.append(("reference precompute_and_cache_latents", time.perf_counter() - t0))
        ref_files = T.snapshot(cdir)
        shutil.rmtree(cdir)
    t0 = time.perf_counter()
    cb.build_cache(cfg, *T.models(), "cpu")
    rows.append(("cache_builder.build_cache", time.perf_counter() - t0))
    if ref_shim.reference_available():
        ours = T.snapshot(cdir)
        assert sorted(ours) == sorted(ref_files)
        for name in ours:
            if name != "dataset_index.pt":
                T.same(ours[name], ref_files[name], name)
    print(f"# caching pass over {n} JPEG images (1.2-1.6 MP) + captions, stand-in models, CPU ({os.cpu_count()} cores), identical output files")
    for name, s in rows:
        print(f"{name:42s} {s:7.2f} s   {n / s:6.1f} images/s")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
