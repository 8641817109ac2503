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


def main():
    tr = ref_shim.import_reference_train()
    gold = {"generator": "tests/golden/make_data_golden.py"}
    for name, ov in VARIANTS.items():
        with tempfile.TemporaryDirectory() as root:
            gold[name] = run(tr.ImageTextLatentDataset, tr.custom_collate_fn, tr.build_epoch_shuffle_batch_schedule,
                             tr.pack_sdxl_sample_schedule, root, ov)
    with open(os.path.join(HERE, "data_golden.json"), "w") as f:
        json.dump(gold, f, indent=0)
    print({k: v["len"] for k, v in gold.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
