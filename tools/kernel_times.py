"""Per-kernel device times (CUPTI through torch.profiler) of selected ops at SDXL shapes: separates a GEMM from the
split-K reduce / tail fix-up launched by the same C-ABI call.   python tools/kernel_times.py"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from aozora_sdxl_training_b200 import _lib, ops  # noqa: E402

BF = torch.bfloat16


def prof(name, fn, flops=None, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in p.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg.setdefault(ev.name[:60], []).append(ev.device_time)
    tot = sum(sum(v) for v in agg.values()) / n
    line = f"{name:28s} total {tot:8.1f} us" + (f" {flops / tot / 1e6:7.1f} TF" if flops else "")
    for k, v in agg.items():
        line += f" | {k.split('(')[0].replace('aoz::', '').replace('void ', '')[:28]} x{len(v) // n} {sum(v) / n:7.1f}"
    print(line, flush=True)


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only == "norms":
        return norms()
    for name, NB, H, Cin, Cout in [("conv_1280", 4, 32, 1280, 1280), ("conv_640", 4, 64, 640, 640), ("conv_320", 4, 128, 320, 320),
                                   ("conv_2560_1280", 4, 32, 2560, 1280), ("conv_960_320", 4, 128, 960, 320)]:
        x = torch.randn(NB, H, H, Cin, device="cuda").to(BF)
        w = (torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02).to(BF)
        dy = torch.randn(NB, H, H, Cout, device="cuda").to(BF)
        wf, wd = ops.pack_conv_weight(w)
        y = torch.empty(NB, H, H, Cout, device="cuda", dtype=BF)
        gw = torch.empty_like(w)
        f = 2.0 * NB * H * H * Cout * Cin * 9
        prof(f"{name}_fwd", lambda: ops.conv_fwd(x, wf, Cout, 3, out=y), f)
        for s in (None, 1, 2, 4):
            prof(f"{name}_wgrad_s{s}", lambda: ops.conv_wgrad(dy, x, 3, grad_w=gw, splits=s), f)
    for name, M, N, K in [("proj_1280", 4096, 1280, 1280), ("qkv_1280", 4096, 3840, 1280), ("ff2_1280", 4096, 1280, 5120),
                          ("ff1_1280", 4096, 10240, 1280), ("proj_640", 16384, 640, 640), ("ff1_640", 16384, 5120, 640)]:
        x = torch.randn(M, K, device="cuda").to(BF)
        w = (torch.randn(N, K, device="cuda") * 0.02).to(BF)
        dy = torch.randn(M, N, device="cuda").to(BF)
        b = torch.zeros(N, device="cuda", dtype=BF)
        f = 2.0 * M * N * K
        o1, o2, o3 = torch.empty(M, N, device="cuda", dtype=BF), torch.empty(M, K, device="cuda", dtype=BF), torch.empty(N, K, device="cuda", dtype=BF)
        prof(f"{name}_fwd", lambda: ops.gemm(x, w, bias=b, out=o1), f)
        prof(f"{name}_dgrad", lambda: ops.gemm(dy, w, b_mn=True, out=o2, splits=1), f)
        prof(f"{name}_wgrad", lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=o3), f)
        prof(f"{name}_colsum", lambda: ops.colsum(dy))
    for name, M, C in [("geglu_1280", 4096, 1280), ("geglu_640", 16384, 640)]:
        x = torch.randn(M, C, device="cuda").to(BF)
        w = (torch.randn(8 * C, C, device="cuda") * 0.02).to(BF)
        b = torch.zeros(8 * C, device="cuda", dtype=BF)
        aux = torch.empty(M, 8 * C, device="cuda", dtype=BF)
        out = torch.empty(M, 4 * C, device="cuda", dtype=BF)
        prof(name, lambda: ops.gemm(x, w, bias=b, epi=ops.EPI_GEGLU, aux=aux, out=out), 2.0 * M * 8 * C * C)
        dyy = torch.randn(M, 4 * C, device="cuda").to(BF)
        prof(name + "_bwd_elementwise", lambda: ops.geglu_bwd(dyy, aux))
    for name, B, H, T, Tk in [("self_4096", 4, 10, 4096, 4096), ("self_1024", 4, 20, 1024, 1024), ("cross_4096", 4, 10, 4096, 77),
                              ("cross_1024", 4, 20, 1024, 77)]:
        q, do = [torch.randn(B, T, H, 64, device="cuda").to(BF) for _ in range(2)]
        k, v = [torch.randn(B, Tk, H, 64, device="cuda").to(BF) for _ in range(2)]
        f = 4.0 * B * H * T * Tk * 64
        prof(f"attn_{name}_fwd", lambda: ops.attn_fwd(q, k, v, 0.125), f)
        o, lse = ops.attn_fwd(q, k, v, 0.125)
        prof(f"attn_{name}_bwd", lambda: ops.attn_bwd(q, k, v, o, do, lse, 0.125), 2.5 * f)
    norms()


def norms():
    for rows, C in [(4096, 1280), (16384, 640)]:
        x = torch.randn(rows, C, device="cuda").to(BF)
        g, bb = torch.ones(C, device="cuda", dtype=BF), torch.zeros(C, device="cuda", dtype=BF)
        y, mean, rstd = ops.layernorm_fwd(x, g, bb)
        prof(f"ln_fwd_{C}", lambda: ops.layernorm_fwd(x, g, bb))
        prof(f"ln_bwd_{C}", lambda: ops.layernorm_bwd(x, x, g, mean, rstd, dres=x))
        dy, dres, col = torch.randn_like(x), torch.randn_like(x), torch.empty(C, device="cuda", dtype=BF)
        prof(f"ln_bwd_colsum_{C}", lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dres=dres, dx_colsum=col))
    for M, half in [(4096, 5120), (16384, 2560)]:
        dyy = torch.randn(M, half, device="cuda").to(BF)
        aux = torch.randn(M, 2 * half, device="cuda").to(BF)
        prof(f"geglu_bwd_{M}x{half}", lambda: ops.geglu_bwd(dyy, aux))
    for NB, HW, C in [(4, 1024, 1280), (4, 4096, 640), (4, 16384, 320), (4, 16384, 960), (4, 1024, 2560)]:
        x = torch.randn(NB, HW, C, device="cuda").to(BF)
        g, bb = torch.ones(C, device="cuda", dtype=BF), torch.zeros(C, device="cuda", dtype=BF)
        y, mean, rstd = ops.groupnorm_fwd(x, g, bb, 1e-5, True)
        prof(f"gn_silu_fwd_{C}x{HW}", lambda: ops.groupnorm_fwd(x, g, bb, 1e-5, True))
        prof(f"gn_silu_bwd_{C}x{HW}", lambda: ops.groupnorm_bwd(x, x, g, bb, mean, rstd, True, dres=x))


if __name__ == "__main__":
    main()
