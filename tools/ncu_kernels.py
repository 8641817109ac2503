"""One launch of every hot kernel family at its SDXL shape, for ONE ncu capture (profiles/rNN_kernels_ncu.txt):

    ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
        --metrics $(python tools/ncu_summary.py --metrics) --clock-control none --profile-from-start off \
        -o gpurun_out/kernels python tools/ncu_kernels.py
    python tools/ncu_summary.py gpurun_out/kernels.ncu-rep > profiles/r01_kernels_ncu_v1.txt

Each op is warmed up outside the profiled range (cudaProfilerStart/Stop bracket exactly one call), so the report holds one
launch per kernel: attention forward / dK,dV / dQ (self 4096 tokens x 10 heads, cross 77 keys), the 3x3
implicit-GEMM convolution (1280 -> 1280 at 32x32: forward, dgrad, wgrad), the GEGLU projection GEMM, GroupNorm+SiLU
(1280 channels, forward and backward), LayerNorm (forward and backward), the timestep-embedding MLP (M = batch), the
weighted-MSE loss, the noising kernel, the gradient-norm reduction and the Raven update."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from aozora_sdxl_training_b200 import ops  # noqa: E402

BF = torch.bfloat16
dev = "cuda"


def once(fn, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def main():
    B = 4
    # attention
    for H, T, Tk in [(10, 4096, 4096), (10, 4096, 77)]:
        q, do = [torch.randn(B, T, H, 64, device=dev).to(BF) for _ in range(2)]
        k, v = [torch.randn(B, Tk, H, 64, device=dev).to(BF) for _ in range(2)]
        o, lse = ops.attn_fwd(q, k, v, 0.125)
        once(lambda: ops.attn_fwd(q, k, v, 0.125))
        once(lambda: ops.attn_bwd(q, k, v, o, do, lse, 0.125))
    # 3x3 convolution as implicit GEMM, 1280 -> 1280 at 32x32
    x = torch.randn(B, 32, 32, 1280, device=dev).to(BF)
    w = (torch.randn(1280, 1280, 3, 3, device=dev) * 0.02).to(BF)
    dy = torch.randn(B, 32, 32, 1280, device=dev).to(BF)
    wf, wd = ops.pack_conv_weight(w)
    y = torch.empty_like(x)
    gw = torch.empty_like(w)
    once(lambda: ops.conv_fwd(x, wf, 1280, 3, out=y))
    once(lambda: ops.conv_fwd(dy, wd, 1280, 3, flip=True, out=y))
    once(lambda: ops.conv_wgrad(dy, x, 3, grad_w=gw))
    # GEGLU projection
    M, C = 4096, 1280
    xa = torch.randn(M, C, device=dev).to(BF)
    wg = (torch.randn(8 * C, C, device=dev) * 0.02).to(BF)
    bg = torch.zeros(8 * C, device=dev, dtype=BF)
    aux = torch.empty(M, 8 * C, device=dev, dtype=BF)
    og = torch.empty(M, 4 * C, device=dev, dtype=BF)
    once(lambda: ops.gemm(xa, wg, bias=bg, epi=ops.EPI_GEGLU, aux=aux, out=og))
    dg = torch.randn(M, 4 * C, device=dev).to(BF)
    once(lambda: ops.geglu_bwd(dg, aux))
    # GroupNorm + SiLU, 1280 channels at 32x32 (BASELINE config 5) and 320 channels at 128x128
    for HW, Cg in [(1024, 1280), (16384, 320)][:int(os.environ.get('NCU_GN_SHAPES', '2'))]:
        xg = torch.randn(B, HW, Cg, device=dev).to(BF)
        ga, be = torch.ones(Cg, device=dev, dtype=BF), torch.zeros(Cg, device=dev, dtype=BF)
        yg, mean, rstd = ops.groupnorm_fwd(xg, ga, be, 1e-5, True)
        once(lambda: ops.groupnorm_fwd(xg, ga, be, 1e-5, True))
        once(lambda: ops.groupnorm_bwd(xg, xg, ga, be, mean, rstd, True, dres=xg))
    # LayerNorm
    for rows, Cl in [(4096, 1280)]:
        xl = torch.randn(rows, Cl, device=dev).to(BF)
        ga, be = torch.ones(Cl, device=dev, dtype=BF), torch.zeros(Cl, device=dev, dtype=BF)
        yl, mean, rstd = ops.layernorm_fwd(xl, ga, be)
        once(lambda: ops.layernorm_fwd(xl, ga, be))
        once(lambda: ops.layernorm_bwd(xl, xl, ga, mean, rstd, dres=xl))
    # timestep-embedding MLP: M = batch rows (weight-read bound)
    te = torch.randn(B, 320, device=dev).to(BF)
    w1 = (torch.randn(1280, 320, device=dev) * 0.02).to(BF)
    b1 = torch.zeros(1280, device=dev, dtype=BF)
    once(lambda: ops.gemm(te, w1, bias=b1))
    # loss + noising
    lat = torch.randn(B, 4, 128, 128, device=dev).to(BF)
    noise = torch.randn(B, 4, 128, 128, device=dev)
    tickets = torch.randint(0, 1000, (B,), device=dev)
    acp = torch.linspace(0.999, 0.005, 1000, device=dev)
    once(lambda: ops.noise_target(lat, noise, tickets, acp, None, "v_prediction"))
    pred = torch.randn(B, 128, 128, 8, device=dev).to(BF)
    table = torch.ones(1000, device=dev)
    once(lambda: ops.mse_loss(pred, lat.float(), tickets, table, float(B)))
    # Raven step + gradient norm over 2^27 parameters (bf16 p, g, m, v: 14 B / parameter -> 1.9 GB per launch)
    from aozora_sdxl_training_b200.optimizers import RavenAdamW
    n = 1 << 21
    params = [torch.nn.Parameter(torch.randn(n, device=dev).to(BF)) for _ in range(64)]
    for p in params:
        p.grad = torch.randn_like(p) * 1e-3
    opt = RavenAdamW(params, lr=1e-5, momentum_dtype=torch.bfloat16)
    once(lambda: opt.step())
    print("ok")


if __name__ == "__main__":
    main()
