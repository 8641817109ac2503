"""Where a kernel's warps spend their time, from an `ncu --set full --import-source on` report: opcode histogram, stall totals, the
hottest SASS lines with their stall reasons, and (optionally) every line of an address window.
    python tools/ncu_hot.py report.ncu-rep [--top 40] [--grep UTCHMMA,SYNCS]"""
import argparse
import csv
import subprocess
from collections import Counter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--grep", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = dict(zip(rows[0], rows[2]))
    for k in ("gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "launch__registers_per_thread"):
        print(f"{k:90s} {d.get(k)}")
    src = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    num = lambda r, h: int(r[ix[h]] or 0)
    tot_i = sum(num(r, "Instructions Executed") for r in data)
    tot_s = sum(num(r, "# Samples") for r in data)
    print("instructions", tot_i, "samples", tot_s)
    ci, cs = Counter(), Counter()
    for r in data:
        op = r[ix["Source"]].split()
        if not op:
            continue
        o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        ci[o] += num(r, "Instructions Executed")
        cs[o] += num(r, "# Samples")
    for o, n in ci.most_common(22):
        print(f"  {o:12s} {n:12d} {100 * n / tot_i:6.2f} % of instructions   {100 * cs[o] / max(tot_s, 1):6.2f} % of samples")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tt = {h[6:]: sum(num(r, h) for r in data) for h in stalls}
    print("stalls:", {k: v for k, v in sorted(tt.items(), key=lambda kv: -kv[1]) if v})
    pats = [p for p in args.grep.split(",") if p]
    top = set(sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:args.top])
    for i, r in enumerate(data):
        if i in top or any(p in r[ix["Source"]] for p in pats):
            st = {h[6:]: num(r, h) for h in stalls if num(r, h)}
            print(f"{i:5d} {r[ix['Source']][:64]:64s} {num(r, '# Samples'):6d} {num(r, 'Instructions Executed'):9d} {st}")


if __name__ == "__main__":
    main()
