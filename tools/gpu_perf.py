"""Kernel microbenchmarks at SDXL shapes (B=4, 1024x1024): TFLOP/s of the GEMM / conv / attention kernels, CTA-pair vs
single-CTA tiles.  Writes gpurun_out/perf.json.   python tools/gpu_perf.py [--quick]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from aozora_sdxl_training_b200 import _lib, ops  # noqa: E402

BF = torch.bfloat16


def timeit(fn, n=10, flush=None):
    """Median device time (ms).  Launches are queued back to back (events around each launch, one synchronize at the
    end) so the GPU never idles waiting for the host inside a timed region; ``flush`` (a buffer > L2) is zeroed first."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    flush = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    res = {}
    lin = [("attn_proj_1280", 4096, 1280, 1280), ("attn_proj_640", 16384, 640, 640), ("ff2_1280", 4096, 1280, 5120),
           ("ff2_640", 16384, 640, 2560), ("ff1_plain_1280", 4096, 10240, 1280), ("conv1x1_320", 65536, 320, 960)]
    for pair in (1,):
        _lib.call("aoz_gemm_set_pair_mode", pair)
        tag = "pair" if pair else "single"
        for name, M, N, K in lin:
            x = torch.randn(M, K, device="cuda").to(BF)
            w = torch.randn(N, K, device="cuda").to(BF)
            dy = torch.randn(M, N, device="cuda").to(BF)
            f = 2.0 * M * N * K
            ms = timeit(lambda: ops.gemm(x, w, splits=1), flush=flush)
            res[f"{name}_fwd_{tag}"] = round(f / ms / 1e9, 1)
            ms = timeit(lambda: ops.gemm(dy, w, b_mn=True, splits=1), flush=flush)
            res[f"{name}_dgrad_{tag}"] = round(f / ms / 1e9, 1)
            ms = timeit(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True), flush=flush)
            res[f"{name}_wgrad_{tag}"] = round(f / ms / 1e9, 1)
        for name, M, C in [("geglu_1280", 4096, 1280), ("geglu_640", 16384, 640)]:
            x = torch.randn(M, C, device="cuda").to(BF)
            w = (torch.randn(8 * C, C, device="cuda") * 0.02).to(BF)
            b = torch.zeros(8 * C, device="cuda", dtype=BF)
            aux = torch.empty(M, 8 * C, device="cuda", dtype=BF)
            ms = timeit(lambda: ops.gemm(x, w, bias=b, epi=ops.EPI_GEGLU, aux=aux), flush=flush)
            res[f"{name}_{tag}"] = round(2.0 * M * 8 * C * C / ms / 1e9, 1)
        for name, NB, H, Cin, Cout in [("conv_1280", 4, 32, 1280, 1280), ("conv_640", 4, 64, 640, 640), ("conv_320", 4, 128, 320, 320),
                                       ("conv_2560_1280", 4, 32, 2560, 1280), ("conv_960_320", 4, 128, 960, 320)]:
            x = torch.randn(NB, H, H, Cin, device="cuda").to(BF)
            w = (torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02).to(BF)
            dy = torch.randn(NB, H, H, Cout, device="cuda").to(BF)
            wf, wd = ops.pack_conv_weight(w)
            f = 2.0 * NB * H * H * Cout * Cin * 9
            ms = timeit(lambda: ops.conv_fwd(x, wf, Cout, 3), flush=flush)
            res[f"{name}_fwd_{tag}"] = round(f / ms / 1e9, 1)
            ms = timeit(lambda: ops.conv_fwd(dy, wd, Cin, 3, flip=True), flush=flush)
            res[f"{name}_dgrad_{tag}"] = round(f / ms / 1e9, 1)
            ms = timeit(lambda: ops.conv_wgrad(dy, x, 3), flush=flush)
            res[f"{name}_wgrad_{tag}"] = round(f / ms / 1e9, 1)
        print(tag, json.dumps({k: v for k, v in res.items() if k.endswith(tag)}), flush=True)
    _lib.call("aoz_gemm_set_pair_mode", 1)
    for name, B, H, T, Tk in [("self_4096", 4, 10, 4096, 4096), ("self_1024", 4, 20, 1024, 1024), ("cross_4096", 4, 10, 4096, 77),
                              ("cross_1024", 4, 20, 1024, 77)]:
        q, do = [torch.randn(B, T, H, 64, device="cuda").to(BF) for _ in range(2)]
        k, v = [torch.randn(B, Tk, H, 64, device="cuda").to(BF) for _ in range(2)]
        f = 4.0 * B * H * T * Tk * 64
        ms = timeit(lambda: ops.attn_fwd(q, k, v, 0.125))
        res[f"attn_{name}_fwd"] = round(f / ms / 1e9, 1)
        res[f"attn_{name}_fwd_us"] = round(ms * 1e3, 1)
        o, lse = ops.attn_fwd(q, k, v, 0.125)
        ms = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, 0.125))
        res[f"attn_{name}_bwd"] = round(2.5 * f / ms / 1e9, 1)
        res[f"attn_{name}_bwd_us"] = round(ms * 1e3, 1)
    print(json.dumps({k: v for k, v in res.items() if k.startswith("attn")}))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "perf.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
