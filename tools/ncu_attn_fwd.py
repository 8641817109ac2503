"""One warm-up and one launch of the attention forward at 4096 tokens x 10 heads (B=4) for an ncu capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aozora_sdxl_training_b200 import ops
B, H, T = 4, 10, 4096
q, k, v, do = [torch.randn(B, T, H, 64, device="cuda").to(torch.bfloat16) for _ in range(4)]
for _ in range(3):
    o, lse = ops.attn_fwd(q, k, v, 0.125)
if len(sys.argv) > 1 and sys.argv[1] == "bwd":
    for _ in range(2):
        ops.attn_bwd(q, k, v, o, do, lse, 0.125)
torch.cuda.synchronize()
print("ok")
