"""Freeze the reference's own data path on the synthetic cache of tests/data_fixture.py (build container only).

    python tests/golden/make_data_golden.py        -> tests/golden/data_golden.json

Executes the REFERENCE's ImageTextLatentDataset / custom_collate_fn / BucketBatchSampler /
build_epoch_shuffle_batch_schedule / pack_sdxl_sample_schedule (train.py:461-534, 767-778, 1992-2254) through
oracle/ref_shim.py and stores content digests; tests/test_data.py compares the product against them anywhere."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from data_fixture import DataCfg, build_cache, item_digest  # noqa: E402
from oracle import ref_shim  # noqa: E402

VARIANTS = {
    "dropout_scale_json": dict(),
    "plain_txt": dict(CAPTION_SOURCE_TYPE="txt", UNCONDITIONAL_DROPOUT=False, TEXT_CONDITIONING_SCALE_ENABLED=False),
    "dropout_only_seed7": dict(SEED=7, TEXT_CONDITIONING_SCALE_ENABLED=False, UNCONDITIONAL_DROPOUT_CHANCE=0.6),
    "scale_only_rf": dict(is_rectified_flow=True, UNCONDITIONAL_DROPOUT=False, CAPTION_TAGS_PERCENT=0, CAPTION_NL_PERCENT=100),
}


def run(dataset_cls, collate_fn, schedule_fn, pack_fn, root, overrides):
    rf = overrides.get("is_rectified_flow", False)
    cfg = type("Cfg", (DataCfg,), dict(overrides, INSTANCE_DATASETS=build_cache(root, rf=rf)))
    ds = dataset_cls(cfg)
    out = {"len": len(ds), "bucket_keys": [list(k) for k in ds.bucket_keys],
           "order": [os.path.relpath(it["lat_path"], root).replace("\\", "/") for it in ds.items]}
    sched = {}
    for bs in (1, 2, 4):
        s = schedule_fn(ds, 17, bs, cfg.SEED)
        sched[str(bs)] = s
    out["schedule"] = sched
    packed = pack_fn(sched["4"], 4)
    out["packed"] = packed
    out["items"] = [[item_digest(ds[p], root) for p in batch] for batch in packed[:8]]
    out["batches"] = [item_digest(collate_fn([ds[p] for p in batch]), root) for batch in packed[:8]]
    return out


LN_M05 = [45, 143, 176, 173, 154, 126, 94, 59, 26, 4]              # GUI logit-normal(-0.5, 1) counts
SPREAD_CASES = [dict(n=23, bs=4, steps=30, seed=42, alloc={"bin_size": 100, "counts": LN_M05}),
                dict(n=7, bs=2, steps=40, seed=7, alloc=None),
                dict(n=40, bs=1, steps=77, seed=42, alloc={"bin_size": 250, "counts": [1, 2, 3, 4]}),
                dict(n=101, bs=5, steps=33, seed=3, alloc={"bin_size": 100, "counts": LN_M05})]
BUCKET_GRID = [(w, h, ta, up) for w in (300, 640, 1000, 1024, 1500, 2048, 4000) for h in (256, 700, 1024, 1536, 3000)
               for ta in (None, 896, 1152, 1536) for up in (False, True)]


def spread_keys(n, seed):
    import random
    r = random.Random(seed)
    return [r.choice([(1024, 1024), (896, 1152), (1216, 832)]) for _ in range(n)]


def run_spread(pool_fn, schedule_fn, case):
    keys = spread_keys(case["n"], case["seed"])
    pool, ranges = pool_fn(case["alloc"], case["steps"] * case["bs"], 1000, case["seed"], False)
    return [[int(i) for i in b] for b in schedule_fn(keys, case["steps"], case["bs"], case["seed"], pool, ranges)]


def run_buckets(optimal_fn, multi_fn, ladder_fn):
    return dict(ladders={str(t): [list(b) for b in ladder_fn(t)] for t in (None, 896, 1024, 1152, 1536, 5000)},
                optimal=[list(optimal_fn(w, h, ta, 64, up)) for w, h, ta, up in BUCKET_GRID],
                multi=[[list(b) for b in multi_fn(w, h, ta, up, 2)] for w, h, ta, up in BUCKET_GRID[::7]])


def main():
    tr = ref_shim.import_reference_train()
    gold = {"generator": "tests/golden/make_data_golden.py"}

    class _DS:
        def __init__(self, keys):
            self.bucket_keys = keys

        def __len__(self):
            return len(self.bucket_keys)
    gold["spread"] = [run_spread(tr.build_timestep_ticket_pool,
                                 lambda keys, steps, bs, seed, pool, ranges: tr.build_image_batch_schedule(_DS(keys), steps, bs, seed, pool, ranges, True), c)
                      for c in SPREAD_CASES]
    gold["buckets"] = run_buckets(tr.get_optimal_bucket, tr.get_multi_bucket_resolutions, tr.get_bucket_ladder)
    for name, ov in VARIANTS.items():
        with tempfile.TemporaryDirectory() as root:
            gold[name] = run(tr.ImageTextLatentDataset, tr.custom_collate_fn, tr.build_epoch_shuffle_batch_schedule,
                             tr.pack_sdxl_sample_schedule, root, ov)
    with open(os.path.join(HERE, "data_golden.json"), "w") as f:
        json.dump(gold, f, indent=0)
    print({k: v["len"] for k, v in gold.items() if isinstance(v, dict) and "len" in v})


if __name__ == "__main__":
    main()
