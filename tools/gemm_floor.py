"""Where does the ~540-cycle per-K-iteration floor of the GEMM main loop come from?  Times the kernel with parts of the
protocol disabled (results are garbage in those modes): 1 = epilogue idle, 2 = MMA free-running (no wait for data)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aozora_sdxl_training_b200 import _lib, ops
from tools.gpu_perf import timeit

M, N, K = 8192, 7680, 2560
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
res = {}
_lib.call("aoz_gemm_set_pair_mode", 0)
for flags in (0, 1, 2, 3):
    _lib.call("aoz_gemm_debug_flags", flags)
    for bn in (64, 128, 256):
        _lib.call("aoz_gemm_force_bn", bn)
        ms = timeit(lambda: ops.gemm(x, w, out=out, splits=1), n=5)
        units = (M // 128) * (N // bn)
        rounds = -(-units // 148)
        res[f"flags{flags}_bn{bn}"] = round(ms * 1e6 / (rounds * (K // 64)), 1)
        print(flags, bn, res[f"flags{flags}_bn{bn}"], "ns/k-iter", flush=True)
_lib.call("aoz_gemm_debug_flags", 0)
_lib.call("aoz_gemm_force_bn", 0)
_lib.call("aoz_gemm_set_pair_mode", 1)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_floor.json"), "w"), indent=1)
