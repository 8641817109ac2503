"""How sensitive is a 50-step loss curve to rounding?  CPU only (oracle vs oracle): the fp32 oracle against the same oracle whose
gradients are rounded to bf16 once per step.  Measured (tiny config, epsilon, lr 1e-4, 50 steps): up to 5.5 % at single steps,
0.9 % on average -- the floor under tests/test_gpu_fullsize.py::test_loss_curve_50_steps_vs_oracle's free-running tolerance.

    python tools/loss_curve_sensitivity.py epsilon 1e-4 50
"""
import os
import sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import host_ref
from oracle.scheduler_ref import RefDDPMScheduler
from oracle.train_step_ref import RefRaven, ref_forward_loss
from oracle.unet_ref import RefUNet2DConditionModel, tiny_config as ref_tiny
from aozora_sdxl_training_b200.unet import UNet2DConditionModel, init_weights_, tiny_config
BF16 = torch.bfloat16
mode = sys.argv[1]; lr = float(sys.argv[2]); steps = int(sys.argv[3])
prod = init_weights_(UNet2DConditionModel(tiny_config()), seed=42, std=0.05).to(BF16)
sd = {k: v.float() for k, v in prod.state_dict().items()}
def run(perturb):
    ref = RefUNet2DConditionModel(ref_tiny()); ref.load_state_dict(sd)
    hp = dict(lr=lr, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8, debias_strength=0.3, momentum_dtype=torch.float32)
    ropt = RefRaven(list(ref.parameters()), **hp)
    rs = host_ref.RefTimestepSampler(steps, 2, 42, None, False)
    sch = RefDDPMScheduler(prediction_type=mode)
    batches = []
    for s in range(4):
        g = torch.Generator().manual_seed(100 + s)
        batches.append(dict(latents=(torch.randn(2, 4, 16, 16, generator=g) * 0.8).to(BF16), embeds=torch.randn(2, 77, 128, generator=g).to(BF16).float(),
                            pooled=torch.randn(2, 64, generator=g).to(BF16).float(), time_ids_data=[[1024, 1024, 0, 0, 1024, 1024]] * 2))
    out = []
    for micro in range(1, steps + 1):
        b = batches[micro % 4]
        ts, _ = rs.sample(2)
        loss, *_ = ref_forward_loss(ref, sch, b, prediction_type=mode, timesteps=ts, micro_step=micro, seed=42, compute_dtype=BF16, autocast=False)
        loss.backward()
        if perturb:
            with torch.no_grad():
                for p in ref.parameters():
                    p.grad.copy_(p.grad.to(BF16).float())
        torch.nn.utils.clip_grad_norm_(list(ref.parameters()), 1.0)
        ropt.step(); ropt.zero_grad(set_to_none=True)
        with torch.no_grad():
            for r in ref.parameters():
                r.copy_(r.to(BF16).float())
        out.append(float(loss))
    return torch.tensor(out)
a = run(False); b = run(True)
rel = (a - b).abs() / a
print(mode, lr, "max", rel.max().item(), "mean", rel.mean().item())
print([round(x, 4) for x in rel.tolist()])
print([round(x, 3) for x in a.tolist()])
