"""Fixed cost of one GEMM launch: kernel duration (CUPTI) against K at the projection shape M=4096, N=1280, for forced tile
plans.  The intercept of the line is what every launch pays before / after its main loop (CTA launch with 227 KB of shared
memory, TMEM allocation, descriptor fetch, pipeline fill, last epilogue, teardown); the slope is the per-K-iteration cost.
    python tools/gemm_fixed.py"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from aozora_sdxl_training_b200 import _lib, ops  # noqa: E402

BF = torch.bfloat16


def kernel_us(fn, n=6, interleave=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        for _ in range(n):
            if interleave is not None:
                interleave()
            fn()
        torch.cuda.synchronize()
    agg = collections.defaultdict(list)
    for ev in p.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and "gemm_bf16_kernel" in ev.name:
            agg["gemm"].append(ev.device_time)
    v = sorted(agg["gemm"])
    return v[len(v) // 2]


def main():
    M, N = 4096, 1280
    parts_only = len(sys.argv) > 1 and sys.argv[1] == "parts"
    big = torch.empty(64 << 20, device="cuda", dtype=torch.float32)           # 256 MB: L2 flush between launches
    flush = lambda: big.zero_()
    for label, pair, bn in ([] if parts_only else [("single_bn160", 0, 160), ("pair_bn256", 2, 256), ("single_bn256", 0, 256)]):
        _lib.call("aoz_gemm_set_pair_mode", pair)
        _lib.call("aoz_gemm_force_bn", bn)
        _lib.call("aoz_gemm_set_tail_mode", 0)
        row = {}
        for K in (64, 128, 256, 512, 1280, 2560):
            x = torch.randn(M, K, device="cuda").to(BF)
            w = (torch.randn(N, K, device="cuda") * 0.02).to(BF)
            res = torch.randn(M, N, device="cuda").to(BF)
            bias = torch.zeros(N, device="cuda", dtype=BF)
            out = torch.empty(M, N, device="cuda", dtype=BF)
            row[f"K{K}"] = round(kernel_us(lambda: ops.gemm(x, w, bias=bias, residual=res, out=out)), 1)
            if K in (64, 1280):
                row[f"K{K}_plain"] = round(kernel_us(lambda: ops.gemm(x, w, out=out, splits=1)), 1)
                row[f"K{K}_cold"] = round(kernel_us(lambda: ops.gemm(x, w, bias=bias, residual=res, out=out), interleave=flush), 1)
        print(label, row, flush=True)
    # grid-size dependence of the fixed part: one tile, one wave, two waves (K = 64: a single K iteration)
    _lib.call("aoz_gemm_set_pair_mode", 0)
    _lib.call("aoz_gemm_force_bn", 128)
    row = {}
    for Mx, Nx in [(128, 128), (128 * 37, 512), (128 * 74, 512), (4096, 1280)]:
        x = torch.randn(Mx, 64, device="cuda").to(BF)
        w = (torch.randn(Nx, 64, device="cuda") * 0.02).to(BF)
        out = torch.empty(Mx, Nx, device="cuda", dtype=BF)
        row[f"{Mx}x{Nx}"] = round(kernel_us(lambda: ops.gemm(x, w, out=out, splits=1)), 1)
    print("K64_single_bn128_by_grid", row, flush=True)
    # what the epilogue of a tile is made of: K = 64 (main loop = one iteration), 320 tiles of 128 x 128 = 3 rounds per CTA.
    # flags: 1 = no epilogue at all, 8 = everything but the global stores, 16 = everything but the TMEM read, 24 = staging only
    _lib.call("aoz_gemm_set_tail_mode", 0)
    x = torch.randn(4096, 64, device="cuda").to(BF)
    w = (torch.randn(1280, 64, device="cuda") * 0.02).to(BF)
    out = torch.empty(4096, 1280, device="cuda", dtype=BF)
    row = {}
    for flags in (0, 1, 8, 16, 24):
        _lib.call("aoz_gemm_debug_flags", flags)
        row[f"dbg{flags}"] = round(kernel_us(lambda: ops.gemm(x, w, out=out, splits=1)), 1)
    _lib.call("aoz_gemm_debug_flags", 0)
    print("K64_4096x1280_bn128_epilogue_parts", row, flush=True)
    # clock64 log of CTA 0's first epilogue warp (dbg 32), K = 64 and K = 1280: cycles per phase of each 32-column chunk
    scratch = ops._gemm_scratch[torch.cuda.current_device()]
    for K in (64, 1280):
        xk = torch.randn(4096, K, device="cuda").to(BF)
        wk = (torch.randn(1280, K, device="cuda") * 0.02).to(BF)
        ops.gemm(xk, wk, out=out, splits=1)
        torch.cuda.synchronize()
        scratch[:8 * 256].zero_()
        _lib.call("aoz_gemm_debug_flags", 32)
        ops.gemm(xk, wk, out=out, splits=1)
        torch.cuda.synchronize()
        _lib.call("aoz_gemm_debug_flags", 0)
        log = scratch[:8 * 256].view(torch.int64).cpu().tolist()
        n, ev = log[0], log[1:1 + log[0]]
        t0 = None
        parts = []
        for v in ev:
            if v < 0:
                t0 = -v
                parts.append("| tile:")
                last = t0
            else:
                parts.append(str(v - last))
                last = v
        print(f"K{K}_epilogue_clock_deltas (tmem, stage, stores per chunk)", " ".join(parts), flush=True)
    _lib.call("aoz_gemm_set_pair_mode", 1); _lib.call("aoz_gemm_force_bn", 0); _lib.call("aoz_gemm_set_tail_mode", 1)


if __name__ == "__main__":
    main()
