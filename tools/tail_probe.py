"""Does cutting the last wave along K pay off for the small projection GEMMs?  Forced (pair, bn) x tail mode, warm timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aozora_sdxl_training_b200 import _lib, ops
from tools.gemm_shapes import timeit_queued

BF = torch.bfloat16
for name, M, N, K, b_mn in [("proj_1280_fwd", 4096, 1280, 1280, False), ("proj_1280_dgrad", 4096, 1280, 1280, True),
                            ("proj_640_fwd", 16384, 640, 640, False), ("ff2_1280_fwd", 4096, 1280, 5120, False),
                            ("qkv_1280_fwd", 4096, 3840, 1280, False), ("qkv_1280_dgrad", 4096, 1280, 3840, True)]:
    x = torch.randn(M, K, device="cuda").to(BF)
    w = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.02).to(BF)
    res = torch.randn(M, N, device="cuda").to(BF)
    bias = torch.zeros(N, device="cuda", dtype=BF)
    out = torch.empty(M, N, device="cuda", dtype=BF)
    fn = (lambda: ops.gemm(x, w, b_mn=True, out=out, splits=1)) if b_mn else (lambda: ops.gemm(x, w, bias=bias, residual=res, out=out))
    row = {}
    for pair in (0, 2):
        _lib.call("aoz_gemm_set_pair_mode", pair)
        for bn in ((128, 256) if b_mn else (128, 160, 192, 256)):
            if pair and bn % (128 if b_mn else 64):
                continue
            _lib.call("aoz_gemm_force_bn", bn)
            for tail in (0, 2):
                _lib.call("aoz_gemm_set_tail_mode", tail)
                plan = _lib.query("aoz_gemm_describe_plan", M, N, K, int(b_mn), 1)
                row[f"{'p' if pair else 's'}{bn}{'+t%dx%d' % ((plan // 1000000) % 100, (plan // 10000) % 100) if tail else ''}"] = round(timeit_queued(fn, n=10) * 1e3, 1)
    _lib.call("aoz_gemm_set_pair_mode", 1); _lib.call("aoz_gemm_force_bn", 0); _lib.call("aoz_gemm_set_tail_mode", 1)
    row["auto"] = round(timeit_queued(fn, n=10) * 1e3, 1)
    print(name, row, flush=True)
