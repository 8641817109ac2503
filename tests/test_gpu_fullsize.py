"""Full-size GPU parity: the real SDXL layout and the shapes bench.py times.

* BASELINE config 1 (SDXL UNet, 512x512 = 64x64x4 latents, batch 1, epsilon, Raven with fp32 moments) through
  ``SDXLTrainStep.step`` against ``oracle.train_step_ref`` on the same bf16-rounded weights (train.py:2719-2784):
  tickets bit-exact, loss rel <= 1e-2, per-block taps cos >= 0.999, flat gradient cos >= 0.999.
* Every kernel family at the shapes the benchmark step launches: attention 4096 tok x 10 heads,
  1024 tok x 20 heads, the bucket tails; the SDXL 3x3 convolutions; GEGLU at C = 1280; Raven over one 29.5 M-element tensor
  and over the whole 1680-tensor / 2.567 B-parameter table against ``oracle.host_ref.raven_update_``.
* The fast erf-GELU (common.cuh ``gelu_cdf``) swept over every bf16 value in [-8, 8] against torch's erf GELU.

References are fp32 with TF32 disabled (plain fp32 matmuls / cuDNN fp32), tolerances stated per test.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(autouse=True)
def _exact_fp32_references():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    torch.cuda.empty_cache()


def _ops():
    from aozora_sdxl_training_b200 import ops
    return ops


def gen(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


def cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten().to(a.device), dim=0).item()


def check(got, ref, rel=4e-3, cmin=0.9999):
    got, ref = got.float().flatten(), ref.float().flatten()
    assert not torch.isnan(got).any()
    r = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
    c = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
    assert r <= rel and c >= cmin, (r, c)


# ------------------------------------------------------------------------------------------------------------------
# BASELINE config 1: the full SDXL UNet, one training step, against the oracle
# ------------------------------------------------------------------------------------------------------------------
def _random_init_(model, seed):
    """N(0, 0.02) everywhere, norm scales 1 + N(0, 0.02): every bias and every gamma gets a non-trivial gradient path."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.normal_(0.0, 0.02, generator=g)
            if p.dim() == 1 and name.endswith("weight"):
                p.add_(1.0)
    return model


def _full_sdxl_step_vs_oracle(mode, lat_hw, px_hw, exclude):
    """One training step of the FULL SDXL layout (1680 tensors, 2,567,463,684 parameters) through SDXLTrainStep against the oracle's
    micro-step (train.py:2719-2784) on identical bf16-rounded weights, latents, text embeddings, tickets and noise."""
    from aozora_sdxl_training_b200 import host
    from aozora_sdxl_training_b200.optimizers import RavenAdamW
    from aozora_sdxl_training_b200.trainer import SDXLTrainStep
    from aozora_sdxl_training_b200.unet import UNet2DConditionModel, sdxl_config
    from oracle import host_ref
    from oracle.scheduler_ref import RefDDPMScheduler
    from oracle.train_step_ref import RefRaven, ref_forward_loss
    from oracle.unet_ref import RefUNet2DConditionModel, sdxl_config as ref_sdxl

    torch.set_num_threads(max(torch.get_num_threads(), min(64, __import__("os").cpu_count() or 1)))
    with torch.device("cuda"):
        prod = UNet2DConditionModel(sdxl_config()).to(BF16)
    _random_init_(prod, 1234)
    assert sum(p.numel() for p in prod.parameters()) == 2_567_463_684 and len(list(prod.parameters())) == 1680
    ref = RefUNet2DConditionModel(ref_sdxl())
    with torch.no_grad():
        for (n, p), (rn, r) in zip(prod.named_parameters(), ref.named_parameters()):
            assert n == rn and p.shape == r.shape, (n, rn)
            r.copy_(p.detach().float().cpu())                       # identical (bf16-rounded) weights
    if exclude:                                                     # train.py:2664-2667 on both sides
        frozen = host.apply_exclusion(prod, exclude)
        ref_frozen = 0
        for n, r in ref.named_parameters():                         # the oracle's own reading of the keywords
            r.requires_grad = not host_ref.is_excluded(n, exclude)
            ref_frozen += 0 if r.requires_grad else r.numel()
        assert ref_frozen == frozen
        assert frozen == 551_102_400 and sum(1 for p in prod.parameters() if not p.requires_grad) == 372
    before = {n: p.detach().clone() for n, p in prod.named_parameters() if not p.requires_grad}

    class Cfg:
        SEED = 42
        BATCH_SIZE = 1
        MAX_TRAIN_STEPS = 10
        GRADIENT_ACCUMULATION_STEPS = 1
        CLIP_GRAD_NORM = 1.0
        PREDICTION_TYPE = mode
        TIMESTEP_ALLOCATION = None
        TIMESTEP_STRATIFIED_SAMPLING = False
        TIMESTEP_LOSS_WEIGHT_CURVE = None
        LR_CUSTOM_CURVE = [[0.0, 1e-4], [1.0, 1e-4]]

    hp = dict(lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8, debias_strength=0.3, momentum_dtype=torch.float32)
    opt = RavenAdamW([{"params": [p for p in prod.parameters() if p.requires_grad], "lr_scale": 1.0}], **hp)
    ropt = RefRaven([r for r in ref.parameters() if r.requires_grad], **hp)
    step = SDXLTrainStep(prod, opt, Cfg)
    g = torch.Generator().manual_seed(7)
    (lh, lw), (ph, pw) = lat_hw, px_hw
    batch = dict(latents=(torch.randn(1, 4, lh, lw, generator=g) * 0.8).to(BF16), embeds=torch.randn(1, 77, 2048, generator=g).to(BF16),
                 pooled=torch.randn(1, 1280, generator=g).to(BF16), time_ids=[[ph, pw, 0, 0, ph, pw]])
    noise = host_ref.step_noise(batch["latents"].shape, Cfg.SEED, 1)
    taps = {}
    res = step.step(batch, noise=noise, taps=taps)

    ts, _ = host_ref.RefTimestepSampler(Cfg.MAX_TRAIN_STEPS, 1, Cfg.SEED, None, False).sample(1)
    assert res.timesteps.cpu().tolist() == ts.tolist()                                   # tickets: bit-exact
    rb = dict(latents=batch["latents"], embeds=batch["embeds"].float(), pooled=batch["pooled"].float(), time_ids_data=batch["time_ids"])
    rtaps = {}
    rloss, rpred, _, _ = ref_forward_loss(ref, RefDDPMScheduler(prediction_type=mode), rb, prediction_type=mode, timesteps=ts,
                                          micro_step=1, seed=Cfg.SEED, compute_dtype=BF16, autocast=False, taps=rtaps)
    rloss.backward()
    rtrain = [r for r in ref.parameters() if r.requires_grad]
    assert all(r.grad is None for r in ref.parameters() if not r.requires_grad)
    rnorm = float(torch.nn.utils.clip_grad_norm_(rtrain, 1.0))
    ropt.step()

    loss = res.loss_value()
    assert abs(loss - float(rloss)) <= 1e-2 * abs(float(rloss)), (loss, float(rloss))     # north_star: loss rel <= 1e-2
    assert set(taps) == set(rtaps) and len(taps) == 7
    for k in rtaps:                                                                       # per-block outputs: cos >= 0.999
        c = cos(taps[k].permute(0, 3, 1, 2), rtaps[k])
        assert c >= 0.999, (k, c)
    assert abs(res.grad_norm_value() - rnorm) <= 4e-2 * rnorm, (res.grad_norm_value(), rnorm)
    # gradients: after step 1 the fp32 first moment is (1 - beta1) * clip_coef * g, so its direction IS the gradient's
    per = []
    for (name, p), r in zip(prod.named_parameters(), ref.parameters()):
        if not p.requires_grad:
            assert p not in opt.state and torch.equal(p.detach(), before[name])           # frozen: no state, not touched
            continue
        m = opt.state[p]["exp_avg"].double().flatten()
        rm = ropt.state[r]["exp_avg"].double().flatten().cuda()
        per.append((name, float(m @ rm), float(m @ m), float(rm @ rm)))
    dots, norms_p, norms_r = (sum(t[i] for t in per) for i in (1, 2, 3))
    flat = dots / math.sqrt(norms_p * norms_r)
    assert flat >= 0.999, flat                                                            # north_star: gradient cos >= 0.999
    # no single tensor that carries a visible share of the gradient is off (tiny-norm tensors are bf16 noise)
    worst = min(((d / math.sqrt(a * b), n) for n, d, a, b in per if b >= 1e-6 * norms_r), default=(1.0, None))
    assert worst[0] >= 0.99, worst
    # second moments are (1 - beta2) * g^2: magnitudes agree too (clip coefficient and gradient scale)
    p0 = prod.mid_block.attentions[0].transformer_blocks[0].ff.net[0].proj.weight
    r0 = ref.mid_block.attentions[0].transformer_blocks[0].ff.net[0].proj.weight
    ratio = float(opt.state[p0]["exp_avg"].float().norm()) / float(ropt.state[r0]["exp_avg"].float().norm())
    assert abs(ratio - 1.0) <= 5e-2, ratio


def test_config1_full_sdxl_train_step_vs_oracle():
    """BASELINE config 1: 512 x 512 (latent 64 x 64), batch 1, epsilon, every parameter trained."""
    _full_sdxl_step_vs_oracle("epsilon", (64, 64), (512, 512), None)


def test_config4_layer_exclusion_nonsquare_bucket_full_sdxl_vs_oracle():
    """BASELINE config 4's ingredients at the full layout: exclusion keywords ["down_blocks.0", "attn2"] (372 tensors / 551,102,400
    parameters frozen: no weight gradient, no optimizer state, values untouched) on a non-square bucket -- latent 36 x 28 = the
    896 x 1152 bucket's 144 x 112 latent at a quarter of the side, so the token counts are 1008 / 252 / 63 (partial attention tiles at
    every level) and the oracle still finishes in seconds.  v_prediction, as config 4."""
    _full_sdxl_step_vs_oracle("v_prediction", (36, 28), (288, 224), ["down_blocks.0", "attn2"])


# ------------------------------------------------------------------------------------------------------------------
# attention at the benchmark shapes
# ------------------------------------------------------------------------------------------------------------------
def _sdpa_fp32(q, k, v, scale):
    """[B,T,H,64] bf16 -> fp32 attention with explicit fp32 matmuls (no flash / TF32 path), autograd-enabled leaves."""
    qr, kr, vr = [t.float().permute(0, 2, 1, 3).contiguous().requires_grad_(True) for t in (q, k, v)]
    outs, lses = [], []
    for b in range(q.shape[0]):
        s = torch.matmul(qr[b], kr[b].transpose(-1, -2)) * scale
        lses.append(torch.logsumexp(s, dim=-1))
        outs.append(torch.matmul(torch.softmax(s, dim=-1), vr[b]))
    return torch.stack(outs), torch.stack(lses), (qr, kr, vr)


@pytest.mark.parametrize("B,H,Tq,Tk", [(4, 10, 4096, 4096), (4, 20, 1024, 1024), (1, 10, 4032, 4032), (1, 20, 988, 988),
                                        (4, 10, 4096, 77), (4, 20, 1024, 77), (1, 10, 4032, 77)])
def test_attention_at_benchmark_shapes(B, H, Tq, Tk):
    """BASELINE config 5 (4096 tok x 10 heads x 64; 77-key cross-attention), the 32x32 level (1024 x 20) and the
    896x1152 bucket's tails (4032 = 31.5 tiles, 988 = 7.7 tiles).  bf16 outputs vs fp32: rel <= 6e-3 / 8e-3 (grads)."""
    ops = _ops()
    g = gen(60 + H)
    q = torch.randn(B, Tq, H, 64, device="cuda", generator=g).to(BF16)
    k = torch.randn(B, Tk, H, 64, device="cuda", generator=g).to(BF16)
    v = torch.randn(B, Tk, H, 64, device="cuda", generator=g).to(BF16)
    do = torch.randn(B, Tq, H, 64, device="cuda", generator=g).to(BF16)
    o, lse = ops.attn_fwd(q, k, v, 0.125)
    orf, lse_ref, (qr, kr, vr) = _sdpa_fp32(q, k, v, 0.125)
    check(o, orf.permute(0, 2, 1, 3), rel=6e-3)
    assert (lse - lse_ref).abs().max().item() < 2e-4
    orf.backward(do.float().permute(0, 2, 1, 3))
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, 0.125)
    check(dq, qr.grad.permute(0, 2, 1, 3), rel=8e-3)
    check(dk, kr.grad.permute(0, 2, 1, 3), rel=8e-3)
    check(dv, vr.grad.permute(0, 2, 1, 3), rel=8e-3)


def test_attention_peaked_rows_long_sequence():
    """Rows whose maximum jumps late in the key sequence (forces the lazy-rescale path across many KV tiles)."""
    ops = _ops()
    g = gen(71)
    B, H, T = 1, 4, 4096
    q = (torch.randn(B, T, H, 64, device="cuda", generator=g) * 2.0).to(BF16)
    k = (torch.randn(B, T, H, 64, device="cuda", generator=g) * 0.5).to(BF16)
    k[:, 3000:3100] *= 6.0                                    # a late block of large keys: scores jump by >> 2^8 there
    k[:, 4000:] *= 9.0
    v = torch.randn(B, T, H, 64, device="cuda", generator=g).to(BF16)
    o, lse = ops.attn_fwd(q, k, v, 0.125)
    orf, lse_ref, _ = _sdpa_fp32(q, k, v, 0.125)
    check(o, orf.permute(0, 2, 1, 3), rel=8e-3)
    assert (lse - lse_ref).abs().max().item() < 1e-3 * max(1.0, lse_ref.abs().max().item())


# ------------------------------------------------------------------------------------------------------------------
# convolutions at the SDXL shapes (batch 4, 1024x1024 latents 128x128)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("NB,H,W,Cin,Cout,stride,up", [
    (4, 32, 32, 1280, 1280, 1, False),      # mid / level-2 resnets
    (4, 32, 32, 2560, 1280, 1, False),      # up_blocks.0 resnets after the skip concat
    (4, 128, 128, 320, 320, 1, False),      # level-0 resnets
    (4, 128, 128, 640, 320, 1, False),      # up_blocks.2 after the skip concat
    (4, 64, 64, 640, 640, 2, False),        # downsampler (stride 2)
    (2, 32, 32, 1280, 1280, 1, True),       # up_blocks.0 upsampler: nearest 2x, then 3x3 at 64x64
])
def test_conv_at_sdxl_shapes(NB, H, W, Cin, Cout, stride, up):
    ops = _ops()
    g = gen(80)
    x = torch.randn(NB, H, W, Cin, device="cuda", generator=g).to(BF16)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * 0.02).to(BF16)
    b = torch.randn(Cout, device="cuda", generator=g).to(BF16)
    wf, wd = ops.pack_conv_weight(w)
    xin = ops.upsample2x_fwd(x) if up else x
    y = ops.conv_fwd(xin, wf, Cout, 3, stride=stride, pad=1, bias=b)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().requires_grad_(True)
    xr_in = torch.nn.functional.interpolate(xr, scale_factor=2.0, mode="nearest") if up else xr
    yr = torch.nn.functional.conv2d(xr_in, wr, b.float(), stride=stride, padding=1)
    check(y, yr.permute(0, 2, 3, 1))
    dy = torch.randn(y.shape, device="cuda", generator=g).to(BF16)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    dyi = dy if stride == 1 else ops.zero_insert2x(dy, H, W)
    dx = ops.conv_fwd(dyi, wd, Cin, 3, stride=1, pad=1, flip=True)
    if up:
        check(ops.upsample2x_bwd(dx), xr.grad.permute(0, 2, 3, 1), rel=6e-3)     # two bf16 roundings (dgrad, then the 2x2 sum)
    else:
        check(dx, xr.grad.permute(0, 2, 3, 1))
    check(ops.conv_wgrad(dy, xin, 3, stride=stride, pad=1), wr.grad)


def test_geglu_at_c1280():
    """ff.net.0 of the 1280-channel transformer blocks: M = 4 x 32 x 32 tokens, 1280 -> 2 x 5120 (the roofline kernel's shape)."""
    ops = _ops()
    g = gen(81)
    M, C = 4096, 1280
    x = torch.randn(M, C, device="cuda", generator=g).to(BF16)
    w = (torch.randn(8 * C, C, device="cuda", generator=g) * 0.03).to(BF16)
    b = torch.randn(8 * C, device="cuda", generator=g).to(BF16)
    aux = torch.empty(M, 8 * C, device="cuda", dtype=BF16)
    out = ops.gemm(x, w, bias=b, epi=ops.EPI_GEGLU, aux=aux)
    proj = (x.float() @ w.float().t() + b.float()).to(BF16)
    check(aux, proj)
    h, gate = proj.float().chunk(2, dim=-1)
    check(out, h * torch.nn.functional.gelu(gate).to(BF16).float())
    dy = torch.randn(M, 4 * C, device="cuda", generator=g).to(BF16)
    pr = proj.float().requires_grad_(True)
    hh, gg = pr.chunk(2, dim=-1)
    (hh * torch.nn.functional.gelu(gg)).backward(dy.float())
    daux = ops.geglu_bwd(dy, aux)
    check(daux, pr.grad, rel=8e-3)
    # the down projection (K = 5120) and the three gradient GEMMs of the pair at this size
    w2 = (torch.randn(C, 4 * C, device="cuda", generator=g) * 0.02).to(BF16)
    check(ops.gemm(out, w2), out.float() @ w2.float().t())
    check(ops.gemm(daux, w, b_mn=True), daux.float() @ w.float())
    check(ops.gemm(daux, x, a_mn=True, b_mn=True), daux.float().t() @ x.float())


def test_gelu_cdf_sweep_every_bf16_in_range():
    """common.cuh gelu_cdf (Abramowitz-Stegun erfc form) vs torch's erf GELU for EVERY bf16 gate value in [-8, 8], observed
    through the GEGLU epilogue (h = 1, gate = x) and through geglu_bwd (dGELU/dx).  Tolerance: forward within one bf16 ulp
    of the exactly rounded value and 1e-5 absolute in the far negative tail; derivative within 4e-3 absolute."""
    ops = _ops()
    bits = torch.arange(0, 1 << 16, dtype=torch.int32)
    vals = bits.to(torch.int16).view(BF16)
    vf = vals.float()
    vals = vals[torch.isfinite(vf) & (vf.abs() <= 8.0) & ((vf.abs() >= 2.0 ** -100) | (vf == 0))]      # normal numbers only
    n = vals.numel()                                          # ~ 26 k distinct values
    M = ((n + 127) // 128) * 128
    gate = torch.zeros(M, dtype=BF16)
    gate[:n] = vals
    K, N2 = 64, 128                                           # projection [M, 2 * 64]: h columns then gate columns
    x = torch.zeros(M, K, dtype=BF16)
    x[:, 0] = gate
    w = torch.zeros(N2, K, dtype=BF16)
    w[64:, 0] = 1.0                                           # gate_j = x[:, 0]
    b = torch.zeros(N2, dtype=BF16)
    b[:64] = 1.0                                              # h_j = 1
    x, w, b = x.cuda(), w.cuda(), b.cuda()
    aux = torch.empty(M, N2, device="cuda", dtype=BF16)
    out = ops.gemm(x, w, bias=b, epi=ops.EPI_GEGLU, aux=aux)
    assert torch.equal(aux[:, 64].cpu(), gate) and (aux[:, :64].float() == 1).all()
    gf = gate.float().cuda()
    exact = torch.nn.functional.gelu(gf.double()).float()
    want = exact.to(BF16).float()
    got = out[:, 0].float()
    ulp = torch.maximum(want.abs(), torch.tensor(2.0 ** -126, device="cuda")) * 2.0 ** -7
    err = (got - exact).abs()
    assert bool(((err <= ulp) | (err <= 1e-5)).all()), float((err / ulp).max())
    assert (got != want).float().mean().item() < 0.05          # and it is the correctly rounded value almost everywhere
    dy = torch.ones(M, 64, device="cuda", dtype=BF16)
    daux = ops.geglu_bwd(dy, aux)                              # d/dgate = h * gelu'(gate) = gelu'(gate); d/dh = gelu(gate)
    gd = gf.double().requires_grad_(True)
    torch.nn.functional.gelu(gd).sum().backward()
    assert (daux[:, 64].float() - gd.grad.float()).abs().max().item() <= 6e-3      # half a bf16 ulp at 1.0 is 3.9e-3
    assert (daux[:, 0].float() - exact).abs().max().item() <= 8e-3 * 8


# ------------------------------------------------------------------------------------------------------------------
# Raven at full size
# ------------------------------------------------------------------------------------------------------------------
HP = dict(lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8, debias_strength=0.3)


def assert_close_1e6(got, ref, scale, what):
    """|got - ref| <= 1e-6 x (|ref| + scale) elementwise.  ``scale`` is the magnitude of the TERMS the value was formed from
    (see ``update_terms``): where they cancel, a plain relative test on the small result would ask for more digits than fp32
    arithmetic in ANY operation order holds (the product follows CUDA ATen's order -- FMA, x * (1/c) -- the CPU oracle follows
    CPU ATen's; both are the reference's code, raven.py:126-143, on different devices)."""
    err = (got.float() - ref.float()).abs()
    tol = 1e-6 * (ref.float().abs() + scale.float().abs()) + 1e-12
    bad = err > tol
    assert not bool(bad.any()), (what, int(bad.sum()), float((err / tol).max()))


def update_terms(before, g, m_prev, v_new, step):
    """Magnitude of the terms of one fp32 Raven parameter update (raven.py:126-143):  p' = p wd - s (b1 m + (1 - b1) g) / denom.
    The first moment is a SUM of two terms of either sign; when they cancel (one element in 1e4 of a 29 M-element tensor does,
    to 1e-3 of their size) the update inherits their absolute rounding error, so the 1e-6 is taken relative to
    |p| + s (|b1 m| + |(1 - b1) g|) / denom -- not to the cancelled result.  Reproduced on the CPU alone by evaluating the two
    operation orders (tools/raven_order_sensitivity.py): the same ~3.1 k of 29.5 M elements differ by up to 60 x 1e-6 x |update|."""
    from oracle import host_ref
    s = host_ref.raven_scalars(HP["lr"], HP["betas"], HP["eps"], HP["weight_decay"], HP["debias_strength"], step)
    denom = v_new.float().sqrt() / s["sqrt_bc2"] + HP["eps"]
    return before.float().abs() + s["step_size"] * (s["beta1"] * m_prev.float().abs() + s["one_m_b1"] * g.float().abs()) / denom


@pytest.mark.parametrize("pdt,mdt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_raven_largest_sdxl_tensor(pdt, mdt):
    """up_blocks.0.resnets.0.conv1.weight: 1280 x 2560 x 3 x 3 = 29,491,200 elements in ONE tensor (hundreds of chunks)."""
    from aozora_sdxl_training_b200.optimizers import RavenAdamW
    from oracle import host_ref
    n = 1280 * 2560 * 9
    g = gen(90)
    p = torch.nn.Parameter((torch.randn(n, device="cuda", generator=g) * 0.02).to(pdt))
    rp = p.detach().cpu().clone()
    rm, rv = torch.zeros(n, dtype=mdt), torch.zeros(n, dtype=mdt)
    opt = RavenAdamW([p], momentum_dtype=mdt, **HP)
    for step in (1, 2):
        grad = (torch.randn(n, device="cuda", generator=g) * 1e-2).to(pdt)
        p.grad = grad
        opt.step()
        before, m_prev = rp.clone(), rm.clone()
        host_ref.raven_update_(rp, grad.cpu(), rm, rv, step=step, **HP)
    got, ref = p.detach().cpu().float(), rp.float()
    if pdt == torch.float32:
        assert_close_1e6(got, ref, update_terms(before, grad.cpu(), m_prev, rv, 2), "p")      # north_star: 1e-6 relative
        assert_close_1e6(opt.state[p]["exp_avg"].cpu(), rm, grad.cpu(), "exp_avg")
        assert torch.allclose(opt.state[p]["exp_avg_sq"].cpu(), rv, rtol=1e-6, atol=1e-20)
        assert (~torch.isclose(got, ref, rtol=1e-6, atol=1e-9)).float().mean().item() < 3e-4  # and almost everywhere in the plain sense
    else:
        assert (got != ref).float().mean().item() < 2e-3 and (got - ref).abs().max() <= ref.abs().max() * 2 ** -7
        assert torch.allclose(opt.state[p]["exp_avg"].cpu().float(), rm.float(), rtol=2 ** -7, atol=1e-12)


def test_raven_and_clip_over_the_sdxl_parameter_table():
    """The real ``aoz_raven_step_mt`` plan: 1680 tensors / 2,567,463,684 fp32 parameters in one launch, two steps with the
    fused clip coefficient, against the oracle's per-tensor update (raven.py:122-147) and torch's clip_grad_norm_ formula
    (train.py:2775-2778) streamed tensor by tensor through the CPU.  fp32: 1e-6 relative (north_star)."""
    from aozora_sdxl_training_b200.optimizers import RavenAdamW
    from aozora_sdxl_training_b200.unet import UNet2DConditionModel, sdxl_config
    from oracle import host_ref
    with torch.device("meta"):
        shapes = [tuple(p.shape) for p in UNet2DConditionModel(sdxl_config()).parameters()]
    assert len(shapes) == 1680 and sum(math.prod(s) for s in shapes) == 2_567_463_684
    g = gen(91)
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g) * 0.02) for s in shapes]
    p0 = [p.detach().clone() for p in ps]
    grads = [[torch.randn(s, device="cuda", generator=g) * (3e-4 * (1 + (i % 7))) for i, s in enumerate(shapes)] for _ in range(2)]
    opt = RavenAdamW([{"params": ps, "lr_scale": 1.0}], momentum_dtype=torch.float32, **HP)
    max_norm = 1.0
    norms = []
    for st in range(2):
        for p, gr in zip(ps, grads[st]):
            p.grad = gr
        out = opt.clip_and_step(max_norm)
        norms.append(out.clone())
        assert opt.last_launches == 3                            # sum of squares, finalize, ONE update launch
    torch.cuda.synchronize()
    # oracle: total norm = norm of the per-tensor fp32 norms; coefficient = min(1, max_norm / (total + 1e-6))
    coefs = []
    for st in range(2):
        total = math.sqrt(sum(float(gr.double().pow(2).sum()) for gr in grads[st]))      # float64 ground truth
        got_norm, got_coef = float(norms[st][0]), float(norms[st][1])
        assert abs(got_norm - total) <= 1e-5 * total, (got_norm, total)                   # fp32 summation of 2.6e9 squares
        coef = min(1.0, max_norm / (total + 1e-6))
        assert coef < 1.0 and abs(got_coef - coef) <= 1e-5 * coef, (got_coef, coef)
        coefs.append(norms[st][1].cpu())                         # the oracle update uses the coefficient just verified
    misses = 0
    for i, p in enumerate(ps):
        rp = p0[i].cpu()
        rm, rv = torch.zeros_like(rp), torch.zeros_like(rp)
        for st in range(2):
            gc = grads[st][i].cpu() * coefs[st]                  # clip_grad_norm_ scales the gradients in place, fp32
            before, m_prev = rp.clone(), rm.clone()
            host_ref.raven_update_(rp, gc, rm, rv, step=st + 1, **HP)
        got = p.detach().cpu()
        assert_close_1e6(got, rp, update_terms(before, gc, m_prev, rv, 2), (i, shapes[i], "p"))
        assert_close_1e6(opt.state[p]["exp_avg"].cpu(), rm, gc, (i, shapes[i], "exp_avg"))
        misses += int((~torch.isclose(got, rp, rtol=1e-6, atol=1e-9)).sum())
    # and in the plain sense, isclose(rtol 1e-6) holds for all but the cancelling elements (~1e-4 of them, see update_terms)
    assert misses < 3e-4 * 2_567_463_684, misses


# ------------------------------------------------------------------------------------------------------------------
# loss curves: 50 optimizer steps, all three prediction types, against the oracle
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["epsilon", "v_prediction", "rectified_flow"])
def test_loss_curve_50_steps_vs_oracle(mode):
    """north_star 'matching reference loss curves within tolerance' (the step is train.py:2719-2784): 50 micro-steps = 50 Raven
    steps at lr 1e-4 on a fixed 4-batch set, so the loss really moves (about 1.5 -> 0.4).  Product: bf16 weights, fp32 moments.

    Two oracles run beside it, both in fp32 math on bf16-STORED weights (after every oracle step the weights are rounded to
    bf16 -- exactly what a bf16 parameter tensor keeps of the fp32 update, raven.py:139-147):

    * along the product's trajectory: before every step the oracle's weights are set to the product's current weights, so
      step k compares the SAME function -- loss within 5e-3 relative and gradient norm within 4e-2, all 50 steps;
    * free-running: the oracle trains on its own.  Training at this rate amplifies any perturbation (tools/loss_curve_sensitivity.py:
      rounding the ORACLE'S OWN gradients to bf16 once per step already moves its curve by up to 5.5 % at single steps), so the
      tolerance is CALIBRATED inside the test: a third oracle trains with its gradients perturbed at the north-star parity bar
      (elementwise relative Gaussian noise of 4.5 %, i.e. gradient cosine 0.999 against the clean oracle, fixed seed), and the
      product's curve may leave the clean oracle's by that oracle's own deviation plus three times the measured drift between two
      VALID variants of the product itself (default kernels vs the round-1 attention kernels -- rounding order only, both pass every
      parity test; tools/loss_curve_variants.py, profiles/r02_loss_curve_variants.txt: epsilon 5.5 % mean / 4.6 % worst window,
      v_prediction 0.1 % / 0.2 %, rectified flow 1.3 % / 3.0 %), plus 1 % absolute.  Measured on B200 (epsilon): product vs clean
      oracle 8.0 % mean / 11.7 % worst window, with 7.7e-4 / 4.5e-3 along the trajectory."""
    from aozora_sdxl_training_b200.optimizers import RavenAdamW
    from aozora_sdxl_training_b200.trainer import SDXLTrainStep
    from aozora_sdxl_training_b200.unet import UNet2DConditionModel, init_weights_, tiny_config
    from oracle import host_ref
    from oracle.scheduler_ref import RefDDPMScheduler
    from oracle.train_step_ref import RefRaven, ref_forward_loss, ref_train_step
    from oracle.unet_ref import RefUNet2DConditionModel, tiny_config as ref_tiny

    prod = init_weights_(UNet2DConditionModel(tiny_config()), seed=42, std=0.05).to(BF16)
    sd = {k: v.float() for k, v in prod.state_dict().items()}
    free, forced, noisy = RefUNet2DConditionModel(ref_tiny()), RefUNet2DConditionModel(ref_tiny()), RefUNet2DConditionModel(ref_tiny())
    free.load_state_dict(sd)
    forced.load_state_dict(sd)
    noisy.load_state_dict(sd)
    prod = prod.cuda()
    steps = 50

    class Cfg:
        SEED = 42
        BATCH_SIZE = 2
        MAX_TRAIN_STEPS = steps
        GRADIENT_ACCUMULATION_STEPS = 1
        CLIP_GRAD_NORM = 1.0
        PREDICTION_TYPE = mode
        TIMESTEP_ALLOCATION = None
        TIMESTEP_STRATIFIED_SAMPLING = False
        TIMESTEP_LOSS_WEIGHT_CURVE = None
        LR_CUSTOM_CURVE = [[0.0, 1e-4], [1.0, 1e-4]]

    hp = dict(lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8, debias_strength=0.3, momentum_dtype=torch.float32)
    opt = RavenAdamW([{"params": list(prod.parameters()), "lr_scale": 1.0}], **hp)
    ropt = RefRaven(list(free.parameters()), **hp)
    nopt = RefRaven(list(noisy.parameters()), **hp)
    ngen = torch.Generator().manual_seed(4242)
    step = SDXLTrainStep(prod, opt, Cfg)
    rsampler = host_ref.RefTimestepSampler(steps, 2, Cfg.SEED, None, False)
    sch = RefDDPMScheduler(prediction_type=mode)
    batches = []
    for s in range(4):
        g = torch.Generator().manual_seed(100 + s)
        batches.append(dict(latents=(torch.randn(2, 4, 16, 16, generator=g) * 0.8).to(BF16), embeds=torch.randn(2, 77, 128, generator=g).to(BF16),
                            pooled=torch.randn(2, 64, generator=g).to(BF16), time_ids=[[1024, 1024, 0, 0, 1024, 1024]] * 2))
    got, want_free, want_forced, want_noisy, gn, gn_forced = [], [], [], [], [], []
    for micro in range(1, steps + 1):
        b = batches[micro % 4]
        rb = dict(latents=b["latents"], embeds=b["embeds"].float(), pooled=b["pooled"].float(), time_ids_data=b["time_ids"])
        ts, _ = rsampler.sample(2)
        # oracle on the product's current weights (bf16 values are exact in fp32)
        with torch.no_grad():
            for r, p in zip(forced.parameters(), prod.parameters()):
                r.copy_(p.detach().float().cpu())
        floss, *_ = ref_forward_loss(forced, sch, rb, prediction_type=mode, timesteps=ts, micro_step=micro, seed=Cfg.SEED,
                                     compute_dtype=BF16, autocast=False)
        floss.backward()
        gn_forced.append(float(torch.nn.utils.clip_grad_norm_(list(forced.parameters()), float("inf"))))
        forced.zero_grad(set_to_none=True)
        want_forced.append(float(floss.detach()))
        # the product's step
        noise = host_ref.step_noise(b["latents"].shape, Cfg.SEED, micro)
        jitter = host_ref.rf_jitter(2, Cfg.SEED, micro)
        res = step.step(b, noise=noise, jitter=jitter)
        assert res.timesteps.cpu().tolist() == ts.tolist()
        got.append(res.loss_value())
        gn.append(res.grad_norm_value())
        # the free-running oracle's step
        rres = ref_train_step(free, sch, ropt, rb, prediction_type=mode, timesteps=ts, micro_step=micro, seed=Cfg.SEED,
                              compute_dtype=BF16, autocast=False, clip_grad_norm=1.0)
        with torch.no_grad():
            for r in free.parameters():
                r.copy_(r.to(BF16).float())                      # bf16 parameter storage
        want_free.append(rres["loss"])
        # the calibration oracle: same step, gradients perturbed at the parity bar (cosine 0.999 <=> 4.5 % relative L2 noise)
        nloss, *_ = ref_forward_loss(noisy, sch, rb, prediction_type=mode, timesteps=ts, micro_step=micro, seed=Cfg.SEED,
                                     compute_dtype=BF16, autocast=False)
        nloss.backward()
        with torch.no_grad():
            for r in noisy.parameters():
                r.grad.mul_(1.0 + 0.045 * torch.randn(r.grad.shape, generator=ngen))
        torch.nn.utils.clip_grad_norm_(list(noisy.parameters()), 1.0)
        nopt.step()
        nopt.zero_grad(set_to_none=True)
        with torch.no_grad():
            for r in noisy.parameters():
                r.copy_(r.to(BF16).float())
        want_noisy.append(float(nloss.detach()))
    got_t, free_t, forced_t = torch.tensor(got), torch.tensor(want_free), torch.tensor(want_forced)
    rel_forced = (got_t - forced_t).abs() / forced_t.abs()
    rel_gn = (torch.tensor(gn) - torch.tensor(gn_forced)).abs() / torch.tensor(gn_forced)
    noisy_t = torch.tensor(want_noisy)
    rel_free = (got_t - free_t).abs() / free_t.abs()
    rel_cal = (noisy_t - free_t).abs() / free_t.abs()
    win = (got_t.view(5, 10).mean(1) / free_t.view(5, 10).mean(1) - 1).abs()
    win_cal = (noisy_t.view(5, 10).mean(1) / free_t.view(5, 10).mean(1) - 1).abs()
    print(f"[loss curve {mode}] along trajectory: loss rel max {rel_forced.max():.2e}, grad-norm rel max {rel_gn.max():.2e}; free-running: "
          f"step max {rel_free.max():.3f}, mean {rel_free.mean():.4f}, window max {win.max():.4f} (calibration oracle at cos 0.999: "
          f"step max {rel_cal.max():.3f}, mean {rel_cal.mean():.4f}, window max {win_cal.max():.4f}); loss {got[0]:.3f} -> "
          f"{sum(got[-10:]) / 10:.3f}")
    assert got_t[-10:].mean() < 0.85 * got_t[:3].mean()          # the run really trained (epsilon 1.50 -> 0.53, v 1.23 -> 0.77, RF 2.12 -> 1.56)
    assert rel_forced.max().item() <= 5e-3, rel_forced           # same weights, same batch: the SAME function at all 50 steps
    assert rel_gn.max().item() <= 4e-2, rel_gn
    drift_mean, drift_win = {"epsilon": (0.0548, 0.0462), "v_prediction": (0.0011, 0.0019), "rectified_flow": (0.0128, 0.0297)}[mode]
    assert win.max().item() <= win_cal.max().item() + 3 * drift_win + 1e-2, (win, win_cal)
    assert rel_free.mean().item() <= rel_cal.mean().item() + 3 * drift_mean + 1e-2, (rel_free.mean(), rel_cal.mean())
