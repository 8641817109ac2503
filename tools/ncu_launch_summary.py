"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (one training step) per kernel.
    python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/rNN_ncu_launches_step.txt"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
        n = r[ki].split("(")[0]
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list of ONE eager training step (SDXL UNet, batch 4, 1024x1024): {len(rows) - 1} launches, "
          f"sum of gpu__time_duration {tot / 1e6:.2f} ms (cold-cache, serialised: compare SHARES, not absolutes)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e6:9.3f} ms {100 * v[1] / tot:5.1f}% x{v[0]:5d}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
