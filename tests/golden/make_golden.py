"""Freeze known-answer vectors by EXECUTING the reference's own code (build container only).

Run:  python tests/golden/make_golden.py        (needs /root/reference; writes host_golden.json
                                                  and raven_golden.pt next to this file)

The reference has no tests, so these vectors -- outputs of the reference's
functions imported through oracle/ref_shim.py -- are what pins the oracle and the product's host
logic.  Nothing here is read from /root/reference at test time; only the frozen files travel.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402


def sha_i64(values):
    return hashlib.sha256(np.asarray(values, dtype="<i8").tobytes()).hexdigest()


def main():
    tr = ref_shim.import_reference_train()
    RavenAdamW, TitanAdamW = ref_shim.import_reference_optimizers()
    out = {"generator": "tests/golden/make_golden.py", "torch": torch.__version__, "numpy": np.__version__}

    # ---- ticket pools (train.py:665-685) -------------------------------------------------
    ln_m05 = [45, 143, 176, 173, 154, 126, 94, 59, 26, 4]       # GUI logit-normal(-0.5, 1) counts, SURVEY 8c
    cases = {
        "uniform16": dict(allocation={"bin_size": 100, "counts": [150] * 10}, total=16, seed=42, stratified=False),
        "uniform16_strat": dict(allocation={"bin_size": 100, "counts": [150] * 10}, total=16, seed=42, stratified=True),
        "none16": dict(allocation=None, total=16, seed=42, stratified=False),
        "ln8000": dict(allocation={"bin_size": 100, "counts": ln_m05}, total=8000, seed=42, stratified=False),
        "ln8000_strat": dict(allocation={"bin_size": 100, "counts": ln_m05}, total=8000, seed=42, stratified=True),
        "ln64000": dict(allocation={"bin_size": 100, "counts": ln_m05}, total=64000, seed=42, stratified=False),
        "bin50_seed7": dict(allocation={"bin_size": 50, "counts": list(range(1, 21))}, total=4096, seed=7, stratified=False),
        "bin50_seed7_strat": dict(allocation={"bin_size": 50, "counts": list(range(1, 21))}, total=4096, seed=7, stratified=True),
        "ragged_zero_bins": dict(allocation={"bin_size": 100, "counts": [0, 5, 0, 0, 9, 1, 0, 0, 0, 2]}, total=333, seed=123, stratified=True),
        "empty_total": dict(allocation={"bin_size": 100, "counts": [1] * 10}, total=0, seed=42, stratified=False),
        "seed0": dict(allocation=None, total=100, seed=0, stratified=False),
        "tiny_total3": dict(allocation={"bin_size": 100, "counts": [150] * 10}, total=3, seed=42, stratified=False),
    }
    pools = {}
    for name, c in cases.items():
        pool, ranges = tr.build_timestep_ticket_pool(c["allocation"], c["total"], 1000, c["seed"], c["stratified"])
        pools[name] = dict(case=c, first=[int(x) for x in pool[:32]], n=len(pool), sha256=sha_i64(pool),
                           ranges=[[int(a), int(b)] for a, b in ranges])
    out["ticket_pools"] = pools
    out["scale_counts"] = {
        "a": dict(counts=[3, 0, 7, 11], total=1000, result=tr._scale_timestep_counts([3, 0, 7, 11], 1000)),
        "b": dict(counts=ln_m05, total=64000, result=tr._scale_timestep_counts(ln_m05, 64000)),
        "c": dict(counts=[1, 1, 1], total=2, result=tr._scale_timestep_counts([1, 1, 1], 2)),
    }
    out["gui_logit_normal_counts"] = {          # SURVEY 8c KATs (GUI code needs PyQt6, not importable)
        "mu-0.5_sigma1_total1000": ln_m05,
        "mu0_sigma1_total1000": [14, 69, 116, 144, 157, 157, 144, 116, 69, 14],
    }

    # ---- TimestepSampler pops (train.py:2163-2208) ---------------------------------------
    class Cfg:
        MAX_TRAIN_STEPS = 5
        BATCH_SIZE = 3
        SEED = 42
        is_rectified_flow = False
        TIMESTEP_ALLOCATION = {"bin_size": 100, "counts": ln_m05}
        TIMESTEP_STRATIFIED_SAMPLING = False
    s = tr.TimestepSampler(Cfg, "cpu")
    pops = [s.sample(3)[0].tolist() for _ in range(7)]        # wraps after 5
    out["sampler_pops"] = pops

    # ---- LR curve (train.py:325-359) -----------------------------------------------------
    class _Opt:
        param_groups = [{"lr": 0.0, "lr_scale": 1.0}]
    curve = [[0.0, 0.0], [0.05, 8.0e-7], [0.85, 8.0e-7], [1.0, 1.0e-7]]
    sch = tr.CustomCurveLRScheduler(_Opt, [list(p) for p in curve], 10000)
    lrs = {}
    for st in (0, 1, 250, 499, 500, 5000, 8499, 9000, 9999, 10000):
        sch.step(st)
        lrs[str(st)] = _Opt.param_groups[0]["lr"]
    out["lr_curve"] = dict(curve=curve, total=10000, lr=lrs)

    # ---- loss-weight tables (train.py:2351-2405) -----------------------------------------
    class LC:
        TIMESTEP_LOSS_WEIGHT_CURVE = [[0, 1], [0.49570201, 2], [1, 1]]
    t1 = tr.timestep_loss_curve_from_config(LC, 1000)
    LC.TIMESTEP_LOSS_WEIGHT_CURVE = {"preset": "bell"}
    t2 = tr.timestep_loss_curve_from_config(LC, 1000)
    LC.TIMESTEP_LOSS_WEIGHT_CURVE = [[0.2, 0.5], [0.8, 3.0]]
    t3 = tr.timestep_loss_curve_from_config(LC, 1000)
    out["loss_tables"] = {
        "tri": dict(points=[[0, 1], [0.49570201, 2], [1, 1]], values=t1.tolist()),
        "bell": dict(points={"preset": "bell"}, values=t2.tolist()),
        "inner": dict(points=[[0.2, 0.5], [0.8, 3.0]], values=t3.tolist()),
    }

    # ---- generators (train.py:248-263) ---------------------------------------------------
    g = tr.seeded_torch_generator("cpu", 42, 1, 0x5D1)
    out["rf_jitter_seed42_step1"] = torch.rand(4, generator=g).tolist()
    gen = torch.Generator(device="cpu")
    n = tr.generate_noise(torch.zeros(1, 4, 2, 2), gen, "cpu", step=3, seed=42)
    out["noise_seed42_step3"] = n.flatten().tolist()

    # ---- weighted MSE (train.py:2408-2416) -----------------------------------------------
    gg = torch.Generator().manual_seed(5)
    pred = torch.randn(3, 4, 8, 8, generator=gg).to(torch.bfloat16)
    targ = torch.randn(3, 4, 8, 8, generator=gg)
    ts = torch.tensor([0, 495, 999])
    out["weighted_mse"] = dict(seed=5, value=float(tr.weighted_sdxl_mse_loss(pred, targ, ts, t1)),
                               value_unweighted=float(tr.weighted_sdxl_mse_loss(pred, targ, ts, None)))

    # ---- key map (train.py:2418-2465) ----------------------------------------------------
    from oracle.unet_ref import RefUNet2DConditionModel, sdxl_config
    with torch.device("meta"):
        m = RefUNet2DConditionModel(sdxl_config())
    names = [k for k, _ in m.named_parameters()]
    mapping = tr.get_unet_key_mapping(names)
    out["key_map"] = dict(n=len(names), sha256_names=hashlib.sha256("\n".join(names).encode()).hexdigest(),
                          sha256_ldm=hashlib.sha256("\n".join(mapping[k] for k in names).encode()).hexdigest(),
                          samples={k: mapping[k] for k in names[::97]})
    # digest of the whole {diffusers key -> LDM key} table, in state_dict order (tests/test_checkpoint.py)
    out["unet_key_mapping_sha256"] = hashlib.sha256("\n".join(f"{k} -> {v}" for k, v in mapping.items()).encode()).hexdigest()

    # ---- Raven / Titan trajectories (raven.py:89-149, titan.py:237-296) -------------------
    traj = {}
    hp = dict(lr=8e-7, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8, debias_strength=0.3)
    for tag, pdt, mdt, lr in (("fp32", torch.float32, torch.float32, 8e-7), ("bf16", torch.bfloat16, torch.bfloat16, 8e-7),
                              ("bf16_biglr", torch.bfloat16, torch.bfloat16, 1e-3), ("fp32_mbf16", torch.float32, torch.bfloat16, 1e-4)):
        for cls_name, cls in (("raven", RavenAdamW), ("titan", TitanAdamW)):
            torch.manual_seed(0)
            p = torch.nn.Parameter(torch.randn(257).to(pdt))
            h = dict(hp, lr=lr)
            opt = cls([p], momentum_dtype=mdt, **h)
            gs = []
            for s_ in range(4):
                gr = (torch.randn(257, generator=torch.Generator().manual_seed(100 + s_)) * 1e-3).to(pdt)
                gs.append(gr.clone())
                # titan offloads in a post-accumulate hook; emulate autograd accumulation
                (p * gr).sum().backward()
                opt.step()
                opt.zero_grad(set_to_none=True)
            st = opt.state[p]
            traj[f"{cls_name}_{tag}"] = dict(p0_seed=0, lr=lr, p=p.detach().clone(), m=st["exp_avg"].clone(),
                                             v=st["exp_avg_sq"].clone(), grads=gs, step=st["step"])
            if hasattr(opt, "close"):
                opt.close()
    torch.save(dict(hparams=hp, traj=traj), os.path.join(HERE, "raven_golden.pt"))

    # clip_grad_norm_ dtype behaviour on torch of this image
    gb = [(torch.randn(1000, generator=torch.Generator().manual_seed(9)) * 2).to(torch.bfloat16)]
    pp = torch.nn.Parameter(torch.zeros(1000, dtype=torch.bfloat16))
    pp.grad = gb[0].clone()
    nrm = torch.nn.utils.clip_grad_norm_([pp], 1.0)
    out["clip_bf16"] = dict(norm=float(nrm), norm_dtype=str(nrm.dtype), first=pp.grad[:8].float().tolist())

    with open(os.path.join(HERE, "host_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "host_golden.json"))


if __name__ == "__main__":
    main()
