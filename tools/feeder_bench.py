"""Can the cached-latent data path keep a B200 fed?  One training step at batch 4 takes ~140 ms (28 imgs/s per GPU), so a rank
needs >= 28 samples/s from disk; this times data.BatchFeeder (background thread, collate, time_ids) against building the same
batches inline, on a synthetic cache with real SDXL payload sizes (latents 4x128x128 bf16, embeds 77x2048 bf16, pooled 1280).
    python tools/feeder_bench.py [items] [batches]"""
import os
import shutil
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from aozora_sdxl_training_b200 import data  # noqa: E402


def main():
    n_items = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n_batches = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    tmp = tempfile.mkdtemp(prefix="feeder_bench_")
    root = os.path.join(tmp, "ds")
    cdir = os.path.join(root, ".precomputed_embeddings_cache_standard_sdxl")
    os.makedirs(cdir)
    g = torch.Generator().manual_seed(0)
    files = []
    for i in range(n_items):
        lat, te = os.path.join(cdir, f"i{i:04d}_lat.pt"), os.path.join(cdir, f"i{i:04d}_te.pt")
        torch.save({"latents": torch.randn(4, 128, 128, generator=g).to(torch.bfloat16)}, lat)
        torch.save({"embeds": torch.randn(77, 2048, generator=g).to(torch.bfloat16), "pooled": torch.randn(1, 1280, generator=g).to(torch.bfloat16)}, te)
        files.append({"lat_path": lat, "te_path": te, "original_size": (1024, 1024), "scaled_size": (1024, 1024), "target_size": (1024, 1024),
                      "crop_coords": (0, 0), "relative_path": f"i{i:04d}.png"})
    torch.save({"files": files}, os.path.join(cdir, "dataset_index.pt"))
    cfg = types.SimpleNamespace(SEED=1, is_rectified_flow=False, INSTANCE_DATASETS=[{"path": root, "repeats": 1}], CAPTION_SOURCE_TYPE="txt",
                                UNCONDITIONAL_DROPOUT=False, TEXT_CONDITIONING_SCALE_ENABLED=False)
    ds = data.CachedLatentDataset(cfg)
    sched = data.pack_sample_schedule(data.epoch_shuffle_batch_schedule(ds.bucket_keys, n_batches, 4, 1), 4)
    t0 = time.perf_counter()
    for packed in sched:                                           # inline: what a NUM_WORKERS=0 loop pays on the training thread
        b = data.collate([ds[i] for i in packed])
        b["time_ids"] = data.time_ids_rows(b)
    inline = time.perf_counter() - t0
    feeder = data.BatchFeeder(ds, sched, pin=False)
    t0 = time.perf_counter()
    waited = 0.0
    for _ in feeder:
        t1 = time.perf_counter()
        time.sleep(0.02)                                           # stand-in for a (much shorter than real) GPU step
        waited += time.perf_counter() - t1
    total = time.perf_counter() - t0
    print(f"# {n_batches} batches of 4 from {n_items} cached items ({os.cpu_count()} host cores, page cache warm)")
    print(f"inline build on the consumer thread : {inline / n_batches * 1e3:7.2f} ms per batch  = {4 * n_batches / inline:7.1f} samples/s")
    print(f"BatchFeeder, consumer busy 20 ms/step: {(total - waited) / n_batches * 1e3:7.2f} ms per batch spent waiting for data")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
