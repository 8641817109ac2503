"""GPU bring-up diagnostics: run each kernel family in its own subprocess (a hang or fault in one case must not
take the others down), compare with torch fp32 references, write gpurun_out/diag.json.

    python tools/gpu_diag.py [case ...]        # no args = all cases
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _stats(got, ref):
    import torch
    got = got.float().flatten()
    ref = ref.float().flatten()
    err = (got - ref).abs()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=0).item() if got.numel() else 1.0
    return dict(max_abs=err.max().item(), ref_max=ref.abs().max().item(), cos=cos,
                rel=(err.norm() / (ref.norm() + 1e-30)).item(), nan=bool(torch.isnan(got).any().item()))


def case_gemm(a_mn, b_mn, M, N, K, bias=False, res=False, splits=1):
    import torch
    from aozora_sdxl_training_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g).to(torch.bfloat16)
    bi = torch.randn(N, device="cuda", generator=g).to(torch.bfloat16) if bias else None
    rs = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if res else None
    out = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, bias=bi, residual=rs, splits=splits)
    torch.cuda.synchronize()
    Af = (A.float().t() if a_mn else A.float())
    Bf = (B.float().t() if b_mn else B.float())
    ref = Af @ Bf.t()
    if bias:
        ref = ref + bi.float()
    if res:
        ref = ref.to(torch.bfloat16).float() + rs.float()
    return _stats(out, ref)


def case_geglu(M, C):
    import torch
    from aozora_sdxl_training_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(M, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(8 * C, C, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(8 * C, device="cuda", generator=g).to(torch.bfloat16)
    aux = torch.empty(M, 8 * C, device="cuda", dtype=torch.bfloat16)
    out = ops.gemm(x, w, bias=b, epi=ops.EPI_GEGLU, aux=aux)
    torch.cuda.synchronize()
    proj = (x.float() @ w.float().t() + b.float()).to(torch.bfloat16)
    h, gate = proj.float().chunk(2, dim=-1)
    ref = h * torch.nn.functional.gelu(gate).to(torch.bfloat16).float()
    s = _stats(out, ref)
    s["aux"] = _stats(aux, proj)
    return s


def case_conv(NB, H, W, Cin, Cout, ks, stride):
    import torch
    from aozora_sdxl_training_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    pad = ks // 2
    x = torch.randn(NB, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(Cout, device="cuda", generator=g).to(torch.bfloat16)
    wf, wd = ops.pack_conv_weight(w)
    y = ops.conv_fwd(x, wf, Cout, ks, stride=stride, pad=pad, bias=b)
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, b.float(), stride=stride, padding=pad)
    out = {"fwd": _stats(y, yr.permute(0, 2, 3, 1))}
    dy = torch.randn(y.shape, device="cuda", generator=g).to(torch.bfloat16)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    if stride == 1:
        dx = ops.conv_fwd(dy, wd, Cin, ks, stride=1, pad=pad, flip=True)
    else:
        dyz = ops.zero_insert2x(dy, H, W)
        dx = ops.conv_fwd(dyz, wd, Cin, ks, stride=1, pad=pad, flip=True)
    torch.cuda.synchronize()
    out["dgrad"] = _stats(dx, xr.grad.permute(0, 2, 3, 1))
    dw = ops.conv_wgrad(dy, x, ks, stride=stride, pad=pad)
    torch.cuda.synchronize()
    out["wgrad"] = _stats(dw, wr.grad)
    return out


def case_attn(B, H, Tq, Tk):
    import torch
    from aozora_sdxl_training_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(B, Tq, H, 64, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(B, Tk, H, 64, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(B, Tk, H, 64, device="cuda", generator=g).to(torch.bfloat16)
    do = torch.randn(B, Tq, H, 64, device="cuda", generator=g).to(torch.bfloat16)
    o, lse = ops.attn_fwd(q, k, v, 0.125)
    torch.cuda.synchronize()
    qr, kr, vr = [t.float().permute(0, 2, 1, 3).requires_grad_(True) for t in (q, k, v)]
    orf = torch.nn.functional.scaled_dot_product_attention(qr, kr, vr)
    lse_ref = torch.logsumexp(torch.einsum("bhqd,bhkd->bhqk", qr, kr) * 0.125, dim=-1)
    out = {"fwd": _stats(o, orf.permute(0, 2, 1, 3)), "lse": _stats(lse, lse_ref)}
    orf.backward(do.float().permute(0, 2, 1, 3))
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, 0.125)
    torch.cuda.synchronize()
    out["dq"] = _stats(dq, qr.grad.permute(0, 2, 1, 3))
    out["dk"] = _stats(dk, kr.grad.permute(0, 2, 1, 3))
    out["dv"] = _stats(dv, vr.grad.permute(0, 2, 1, 3))
    return out


def case_norms():
    import torch
    from aozora_sdxl_training_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    out = {}
    for (NB, HW, C, silu) in [(2, 64, 320, True), (2, 256, 1280, False), (1, 100, 960, True)]:
        x = (torch.randn(NB, HW, C, device="cuda", generator=g) * 2 + 0.5).to(torch.bfloat16)
        ga = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).to(torch.bfloat16)
        be = (0.1 * torch.randn(C, device="cuda", generator=g)).to(torch.bfloat16)
        y, mean, rstd = ops.groupnorm_fwd(x, ga, be, 1e-5, silu)
        xr = x.float().requires_grad_(True)
        gr, br = ga.float().requires_grad_(True), be.float().requires_grad_(True)
        yr = torch.nn.functional.group_norm(xr.permute(0, 2, 1), 32, gr, br, 1e-5)
        if silu:
            yr = torch.nn.functional.silu(yr)
        yr = yr.permute(0, 2, 1)
        dy = torch.randn(y.shape, device="cuda", generator=g).to(torch.bfloat16)
        yr.backward(dy.float())
        dx, dg, db = ops.groupnorm_bwd(dy, x, ga, be, mean, rstd, silu)
        torch.cuda.synchronize()
        out[f"gn_{NB}_{HW}_{C}_{int(silu)}"] = dict(fwd=_stats(y, yr), dx=_stats(dx, xr.grad), dgamma=_stats(dg, gr.grad),
                                                    dbeta=_stats(db, br.grad))
    for (rows, C) in [(300, 640), (1000, 1280)]:
        x = (torch.randn(rows, C, device="cuda", generator=g) * 2 + 0.5).to(torch.bfloat16)
        ga = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).to(torch.bfloat16)
        be = (0.1 * torch.randn(C, device="cuda", generator=g)).to(torch.bfloat16)
        y, mean, rstd = ops.layernorm_fwd(x, ga, be)
        xr = x.float().requires_grad_(True)
        gr, br = ga.float().requires_grad_(True), be.float().requires_grad_(True)
        yr = torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-5)
        dy = torch.randn(y.shape, device="cuda", generator=g).to(torch.bfloat16)
        yr.backward(dy.float())
        dx, dg, db = ops.layernorm_bwd(dy, x, ga, mean, rstd)
        torch.cuda.synchronize()
        out[f"ln_{rows}_{C}"] = dict(fwd=_stats(y, yr), dx=_stats(dx, xr.grad), dgamma=_stats(dg, gr.grad), dbeta=_stats(db, br.grad))
    return out


def case_raven():
    import torch
    from aozora_sdxl_training_b200.optimizers import RavenAdamW
    from oracle import host_ref
    out = {}
    for pdt, mdt in [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16)]:
        torch.manual_seed(0)
        shapes = [(257,), (64, 33), (3, 3, 16, 16), (20000,)]
        ps = [torch.nn.Parameter(torch.randn(s).to(pdt).cuda()) for s in shapes]
        ref_p = [p.detach().cpu().clone() for p in ps]
        ref_m = [torch.zeros_like(p, dtype=mdt) for p in ref_p]
        ref_v = [torch.zeros_like(p, dtype=mdt) for p in ref_p]
        hp = dict(lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8, debias_strength=0.3)
        opt = RavenAdamW(ps, momentum_dtype=mdt, **hp)
        for step in range(1, 4):
            for i, p in enumerate(ps):
                g = (torch.randn(p.shape, generator=torch.Generator().manual_seed(100 * step + i)) * 1e-2).to(pdt)
                p.grad = g.cuda()
                host_ref.raven_update_(ref_p[i], g, ref_m[i], ref_v[i], step=step, **hp)
            opt.step()
        torch.cuda.synchronize()
        worst = 0.0
        for i, p in enumerate(ps):
            d = (p.detach().cpu().float() - ref_p[i].float()).abs() / (ref_p[i].float().abs() + 1e-12)
            worst = max(worst, d.max().item())
        mw = max(((opt.state[p]["exp_avg"].cpu().float() - ref_m[i].float()).abs().max().item()) for i, p in enumerate(ps))
        out[f"{pdt}_{mdt}"] = dict(max_rel_p=worst, max_abs_m=mw)
    return out


CASES = {
    "attn_self_small": lambda: case_attn(1, 2, 128, 128),
    "attn_self": lambda: case_attn(2, 5, 1024, 1024),
    "attn_self_ragged": lambda: case_attn(1, 3, 1008, 1008),
    "attn_cross": lambda: case_attn(2, 10, 1024, 77),
    "attn_cross_ragged": lambda: case_attn(1, 5, 988, 154),
    "raven": lambda: case_raven(),
    "norms": lambda: case_norms(),
    "gemm_tn_small": lambda: case_gemm(False, False, 128, 128, 64),
    "gemm_tn": lambda: case_gemm(False, False, 1000, 640, 320),
    "gemm_tn_bn256": lambda: case_gemm(False, False, 4096, 1280, 1280, bias=True, res=True),
    "gemm_tn_tiny_m": lambda: case_gemm(False, False, 4, 1280, 320, bias=True),
    "gemm_a_mn": lambda: case_gemm(True, False, 256, 256, 256),
    "gemm_b_mn": lambda: case_gemm(False, True, 256, 256, 256),
    "gemm_nn_dgrad": lambda: case_gemm(False, True, 1000, 320, 640),
    "gemm_wgrad": lambda: case_gemm(True, True, 640, 320, 1000),
    "gemm_wgrad_split": lambda: case_gemm(True, True, 640, 640, 16384, splits=8),
    "geglu": lambda: case_geglu(512, 128),
    "geglu_big": lambda: case_geglu(4096, 640),
    "conv3_s1": lambda: case_conv(2, 16, 16, 64, 128, 3, 1),
    "conv3_s1_ragged": lambda: case_conv(1, 18, 14, 320, 192, 3, 1),
    "conv3_s2": lambda: case_conv(2, 16, 16, 64, 64, 3, 2),
    "conv1": lambda: case_conv(2, 16, 16, 128, 64, 1, 1),
    "conv_in": lambda: case_conv(2, 16, 16, 8, 320, 3, 1),
    "conv_out": lambda: case_conv(2, 16, 16, 320, 8, 3, 1),
}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        import torch  # noqa
        res = CASES[sys.argv[2]]()
        print("RESULT " + json.dumps(res))
        return
    names = sys.argv[1:] or list(CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = {}
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True, timeout=180)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if r.returncode == 0 and line:
                results[n] = json.loads(line[-1][7:])
            else:
                results[n] = dict(error=(r.stderr or r.stdout)[-1500:], rc=r.returncode)
        except subprocess.TimeoutExpired:
            results[n] = dict(error="TIMEOUT")
        results[n]["_secs"] = round(time.time() - t0, 1) if isinstance(results[n], dict) else None
        print(n, json.dumps(results[n])[:600], flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
