"""Summarise an .ncu-rep (one kernel launch, --set full) into the few numbers the roofline needs.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_ncu.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"kernel: {d.get('Kernel Name')}   id {d.get('ID')}")
        for k in WANT:
            if k in d:
                print(f"  {k:80s} {d[k]:>16s} {units[hdr.index(k)]}")
        try:                                      # derived: achieved DRAM bandwidth of this launch
            def val(k):
                v, u = float(d[k].replace(",", "")), units[hdr.index(k)].lower()
                scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
                         "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0, "s": 1.0}.get(u, 1.0)
                return v * scale
            byts = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            print(f"  {'derived: DRAM read+write / duration':80s} {byts / val('gpu__time_duration.sum') / 1e9:16.1f} GB/s")
        except (KeyError, ValueError, ZeroDivisionError):
            pass
        print()


if __name__ == "__main__":
    if sys.argv[1] == "--metrics":           # the metric list as an `ncu --metrics` argument (lighter than --set full)
        print(",".join(w for w in WANT if not w.startswith("launch__")))
    else:
        main(sys.argv[1])
