"""The bench.py JSON line the driver consumes: checked on the lines recorded from real B200 runs (profiles/) and on the
argument parser, without a GPU.  A missing key would silently void a round's measurement, so the contract is a test."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"}
E2E_KEYS = {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
ROOFLINE_KEYS = {"bound", "achieved", "peak", "unit", "frac", "traffic"}
CPU_KEYS = {"value", "unit", "cores", "kind", "sample"}


def _line(name):
    with open(os.path.join(PROFILES, name)) as f:
        return json.loads([l for l in f.read().splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("name,n", [("r01_bench_v7.json", 1), ("r01_bench_dp2_v2.json", 2), ("r01_bench_dp4_v1.json", 4),
                                    ("r01_bench_dp8_v2.json", 8), ("r02_bench_v3.json", 1), ("r02_bench_dp2_final.json", 2),
                                    ("r02_bench_dp8_mb512.json", 8), ("r02_bench_w3.json", 1), ("r02_bench_w4.json", 1),
                                    ("r02_bench_dp8_w3.json", 8), ("r02_bench_dp8_w4.json", 8), ("r02_bench_dp1_v4.json", 1), ("r02_bench_dp1_v5.json", 1), ("r02_bench_dp2_v2.json", 2),
                                    ("r02_bench_dp8_v2.json", 8)])
def test_recorded_bench_lines_follow_the_contract(name, n):
    d = _line(name)
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["n_gpus"] == n and d["unit"] == "imgs/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None                       # BASELINE.md publishes no number for this metric
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    assert E2E_KEYS <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
    r = d["roofline"]
    assert ROOFLINE_KEYS <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    # whole-job value = images of all ranks / max-over-ranks step time
    imgs_per_step = d["config"]["global_batch"]
    assert abs(d["value"] - imgs_per_step / (d["ms_per_step"] * 1e-3)) / d["value"] < 5e-3
    if n == 1 and d.get("cpu_baseline") is not None:      # (recorded with --no-cpu-baseline: absent)
        assert CPU_KEYS <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("reference", "port")


def test_recorded_reference_arm_line():
    """`bench.py --impl reference` as recorded on the GPU box's host: BASELINE config 1 on all host cores, split timings."""
    d = _line("r02_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["unit"] == "imgs/s" and d["higher_is_better"] is True and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert CPU_KEYS <= set(cb) and cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert abs(cb["seconds_per_step"] - (cb["fwd_bwd_s"] + cb["clip_s"] + cb["raven_s"])) < 0.05
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "config 1" in d["config"]["workload"] and d["config"]["same_config"] is False


def test_bench_cli_defaults_and_reference_arm_flag():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
