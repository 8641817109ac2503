"""A few launches of the GEMM kernel exactly as bench.py's `roofline` times it (GEGLU feed-forward projection,
M=4096 K=1280 N=10240, bias + pre-activation side output), for `ncu --set full`.   python tools/ncu_gemm.py [pair_mode] [plain | proj]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aozora_sdxl_training_b200 import _lib, ops
M, N, K = 4096, 10240, 1280
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
plain = len(sys.argv) > 2 and sys.argv[2] == "plain"
_lib.call("aoz_gemm_set_pair_mode", mode)
proj = len(sys.argv) > 2 and sys.argv[2] == "proj"
if proj:
    # the C x C projection with bias + residual (to_out): the 320-wide one-wave plan and the fast store epilogue
    N = 1280
    w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(x, w, bias=b, residual=res, out=out)
elif plain:
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(x, w, out=out, splits=1)
else:
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N // 2, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(x, w, bias=b, epi=ops.EPI_GEGLU, aux=aux, out=out)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("ok")
