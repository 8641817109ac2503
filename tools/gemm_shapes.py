"""Per-shape tile sweep of the tcgen05 GEMM / implicit-GEMM conv at the SDXL UNet's shapes (B=4, 1024x1024).

For every (shape, op) the kernel is timed with the host cost model's own choice ("auto") and with every forced
(pair mode, tile width) combination; weight-gradient GEMMs are also swept over the split-K factor.  The output
(gpurun_out/gemm_shapes.json) is what the cost model in csrc/gemm.cu is calibrated against.

Timing: 8 back-to-back launches between one CUDA-event pair (warm L2: inside a training step the operands of these
mid-size GEMMs were just written by the previous kernel).

    python tools/gemm_shapes.py [--quick]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from aozora_sdxl_training_b200 import _lib, ops  # noqa: E402

BF = torch.bfloat16
_flush = None


def timeit_queued(fn, n=8):
    """Mean device time (ms) of n back-to-back launches (warm L2, as inside a training step where the operands were just
    produced; one event pair around the whole batch: the 2 us event granularity is amortised)."""
    fn()
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def sweep(fn, b_mn, geglu=False):
    """{config: ms}: auto plus every forced (pair, bn)."""
    out = {}
    _lib.call("aoz_gemm_set_pair_mode", 1)
    _lib.call("aoz_gemm_force_bn", 0)
    out["auto"] = timeit_queued(fn)
    _lib.call("aoz_gemm_set_tail_mode", 0)
    out["auto_notail"] = timeit_queued(fn)
    _lib.call("aoz_gemm_set_tail_mode", 1)
    step = 64 if b_mn else 32
    for pair in (0, 2):
        _lib.call("aoz_gemm_set_pair_mode", pair)
        for bn in range(64, 257, step):
            if geglu and bn not in (128, 256):
                continue
            if pair and bn % (2 * step):
                continue
            _lib.call("aoz_gemm_force_bn", bn)
            try:
                out[f"{'pair' if pair else 'single'}_bn{bn}"] = timeit_queued(fn)
            except Exception as e:           # a forced combination the kernel rejects
                out[f"{'pair' if pair else 'single'}_bn{bn}"] = None
                print("  rejected", pair, bn, str(e)[:80], flush=True)
    _lib.call("aoz_gemm_set_pair_mode", 1)
    _lib.call("aoz_gemm_force_bn", 0)
    return out


def report(name, flops, res, store):
    best = min((v, k) for k, v in res.items() if v)
    store[name] = dict(flops=flops, auto_us=round(res["auto"] * 1e3, 2), auto_tflops=round(flops / res["auto"] / 1e9, 1),
                       best=best[1], best_us=round(best[0] * 1e3, 2), best_tflops=round(flops / best[0] / 1e9, 1),
                       all_us={k: (round(v * 1e3, 2) if v else None) for k, v in res.items()})
    print(f"{name:34s} auto {res['auto'] * 1e3:8.1f} us {flops / res['auto'] / 1e9:7.1f} TF | best {best[1]:14s} "
          f"{best[0] * 1e3:8.1f} us {flops / best[0] / 1e9:7.1f} TF", flush=True)


def main():
    quick = "--quick" in sys.argv
    store = {}
    # (name, M, N, K, count per step) -- Linear layers at B=4, 1024x1024: M = tokens
    lin = [("proj_1280", 4096, 1280, 1280), ("proj_640", 16384, 640, 640), ("qkv_1280", 4096, 3840, 1280),
           ("qkv_640", 16384, 1920, 640), ("ff2_1280", 4096, 1280, 5120), ("ff2_640", 16384, 640, 2560),
           ("ff1_1280", 4096, 10240, 1280), ("ff1_640", 16384, 5120, 640), ("kv_cross_1280", 308, 2560, 2048),
           ("kv_cross_640", 308, 1280, 2048)]
    if quick:
        lin = lin[:2]
    for name, M, N, K in lin:
        x = torch.randn(M, K, device="cuda").to(BF)
        w = (torch.randn(N, K, device="cuda") * 0.02).to(BF)
        dy = torch.randn(M, N, device="cuda").to(BF)
        out_f = torch.empty(M, N, device="cuda", dtype=BF)
        out_d = torch.empty(M, K, device="cuda", dtype=BF)
        out_w = torch.empty(N, K, device="cuda", dtype=BF)
        bias = torch.zeros(N, device="cuda", dtype=BF)
        f = 2.0 * M * N * K
        report(f"{name}_fwd", f, sweep(lambda: ops.gemm(x, w, bias=bias, out=out_f), False), store)
        report(f"{name}_dgrad", f, sweep(lambda: ops.gemm(dy, w, b_mn=True, out=out_d, splits=1), True), store)
        # weight gradient: sweep the split-K factor with the model's tile choice, then tiles at the auto split
        res = {}
        auto_s = _lib.query("aoz_gemm_auto_splits", N, K, M, 1)
        for s in sorted({1, 2, 3, 4, 6, 8, 12, 16, auto_s}):
            if s > (M // 64 + 1) // 2:
                continue
            res[f"splits{s}"] = timeit_queued(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=out_w, splits=s))
        res["auto"] = timeit_queued(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=out_w))
        best_s = int(min((v, k) for k, v in res.items() if k != "auto")[1][6:])
        tiles = sweep(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=out_w, splits=best_s), True)
        for k, v in tiles.items():
            if k != "auto":
                res[f"s{best_s}_{k}"] = v
        store[f"{name}_wgrad_auto_splits"] = auto_s
        report(f"{name}_wgrad", f, res, store)
    for name, M, C in ([] if quick else [("geglu_1280", 4096, 1280), ("geglu_640", 16384, 640)]):
        x = torch.randn(M, C, device="cuda").to(BF)
        w = (torch.randn(8 * C, C, device="cuda") * 0.02).to(BF)
        b = torch.zeros(8 * C, device="cuda", dtype=BF)
        aux = torch.empty(M, 8 * C, device="cuda", dtype=BF)
        out = torch.empty(M, 4 * C, device="cuda", dtype=BF)
        report(name, 2.0 * M * 8 * C * C, sweep(lambda: ops.gemm(x, w, bias=b, epi=ops.EPI_GEGLU, aux=aux, out=out), False, geglu=True), store)
    convs = [("conv_1280", 4, 32, 1280, 1280), ("conv_640", 4, 64, 640, 640), ("conv_320", 4, 128, 320, 320),
             ("conv_2560_1280", 4, 32, 2560, 1280), ("conv_1920_640", 4, 64, 1920, 640), ("conv_960_320", 4, 128, 960, 320),
             ("conv_640_320", 4, 128, 640, 320), ("conv_1280_640", 4, 64, 1280, 640)]
    for name, NB, H, Cin, Cout in ([] if quick else convs):
        x = torch.randn(NB, H, H, Cin, device="cuda").to(BF)
        w = (torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02).to(BF)
        dy = torch.randn(NB, H, H, Cout, device="cuda").to(BF)
        wf, wd = ops.pack_conv_weight(w)
        y = torch.empty(NB, H, H, Cout, device="cuda", dtype=BF)
        dx = torch.empty(NB, H, H, Cin, device="cuda", dtype=BF)
        gw = torch.empty_like(w)
        f = 2.0 * NB * H * H * Cout * Cin * 9
        report(f"{name}_fwd", f, sweep(lambda: ops.conv_fwd(x, wf, Cout, 3, out=y), False), store)
        report(f"{name}_dgrad", f, sweep(lambda: ops.conv_fwd(dy, wd, Cin, 3, flip=True, out=dx), False), store)
        report(f"{name}_wgrad", f, sweep(lambda: ops.conv_wgrad(dy, x, 3, grad_w=gw), True), store)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(store, open(os.path.join(ROOT, "gpurun_out", "gemm_shapes.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
