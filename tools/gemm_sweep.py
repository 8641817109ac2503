"""GEMM tile-width sweep: time per K-iteration vs bn, single CTA vs CTA pair (calibrates the host cost model)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aozora_sdxl_training_b200 import _lib, ops
from tools.gpu_perf import timeit

def main():
    only = sys.argv[1:] and sys.argv[1] == "--one"
    M, N, K = 8192, 7680, 2560        # 64 m-tiles; N divisible by 64..256 multiples of 32? 7680 = 30*256 = 40*192 = 48*160 = 60*128 = 80*96 = 120*64
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = {}
    for pair in ((0,) if only else (0, 2)):
        _lib.call("aoz_gemm_set_pair_mode", pair)
        for bn in ((256,) if only else (64, 96, 128, 160, 192, 224, 256)):
            if pair and bn % 64:
                continue
            _lib.call("aoz_gemm_force_bn", bn)
            ms = timeit(lambda: ops.gemm(x, w, out=out, splits=1), n=(2 if only else 7))
            tf = 2.0 * M * N * K / ms / 1e9
            units = (M // 128) * (N // bn) / (2 if pair else 1)
            rounds = -(-units // (74 if pair else 148))
            ns_per_kiter = ms * 1e6 / (rounds * (K // 64))
            res[f"{'pair' if pair else 'single'}_bn{bn}"] = dict(tflops=round(tf, 1), ms=round(ms, 4), ns_per_kiter=round(ns_per_kiter, 1))
            print(pair, bn, res[f"{'pair' if pair else 'single'}_bn{bn}"], flush=True)
    _lib.call("aoz_gemm_force_bn", 0)
    _lib.call("aoz_gemm_set_pair_mode", 1)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_sweep.json"), "w"), indent=1)

main()
