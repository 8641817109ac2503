"""A few launches of the GEMM kernel at the GEGLU / feed-forward shape for ncu (tools: see profiles/README)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aozora_sdxl_training_b200 import _lib, ops
M, N, K = 4096, 10240, 1280
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.call("aoz_gemm_set_pair_mode", mode)
_lib.call("aoz_gemm_force_bn", 256)
for _ in range(4):
    ops.gemm(x, w, out=out, splits=1)
torch.cuda.synchronize()
print("ok")
