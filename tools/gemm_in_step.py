"""In-step time of every GEMM / conv launch by shape: one eager training step under torch.profiler with ops.trace
recording the call sequence; the i-th gemm_bf16_kernel launch belongs to the i-th recorded call.
    python tools/gemm_in_step.py  -> gpurun_out/gemm_in_step.txt"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import Cfg, synth_batch  # noqa: E402


def main():
    from aozora_sdxl_training_b200 import ops
    from aozora_sdxl_training_b200.optimizers import RavenAdamW
    from aozora_sdxl_training_b200.trainer import SDXLTrainStep
    from aozora_sdxl_training_b200.unet import UNet2DConditionModel, init_weights_fast_, sdxl_config
    dev = torch.device("cuda", 0)
    with torch.device(dev):
        unet = UNet2DConditionModel(sdxl_config()).to(torch.bfloat16)
    init_weights_fast_(unet)
    cfg = type("C", (Cfg,), dict(BATCH_SIZE=4, PREDICTION_TYPE="v_prediction", TIMESTEP_ALLOCATION=None))
    opt = RavenAdamW([{"params": list(unet.parameters()), "lr_scale": 1.0}], lr=8e-7, momentum_dtype=torch.bfloat16, **Cfg.RAVEN)
    step = SDXLTrainStep(unet, opt, cfg, device=dev, use_cuda_graph=False)
    batch = synth_batch(4, 1024, 1, device=dev)
    for _ in range(2):
        step.step(batch)
    torch.cuda.synchronize()
    ops.trace = []
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step.step(batch)
        torch.cuda.synchronize()
    calls, ops.trace = ops.trace, None
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    gemms = [e for e in evs if "gemm_bf16_kernel" in e.name]
    assert len(gemms) == len(calls), (len(gemms), len(calls))
    # attribute the reduce / fix-up kernel that follows a GEMM to the same call
    extra = collections.defaultdict(float)
    gi = -1
    for e in evs:
        if "gemm_bf16_kernel" in e.name:
            gi += 1
        elif gi >= 0 and ("splitk_reduce" in e.name or "tail_fixup" in e.name):
            extra[gi] += e.device_time
    agg = collections.OrderedDict()
    for i, (c, e) in enumerate(zip(calls, gemms)):
        key = c[:8] + ("pair" if "<true>" in e.name or "ILb1" in e.name else "single",)
        a = agg.setdefault(key, [0, 0.0, 0.0, c[8]])
        a[0] += 1
        a[1] += e.device_time
        a[2] += extra.get(i, 0.0)
    rows = sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2]))
    tot = sum(v[1] + v[2] for _, v in rows)
    lines = [f"in-step GEMM/conv time by shape (one eager step, B=4 1024x1024): {tot / 1e3:.2f} ms in {len(calls)} calls",
             f"{'kind':10s} {'M':>6s} {'N':>6s} {'K':>6s} aT bT epi spl tile   count   gemm_us  +red_us  total_ms  TFLOP/s(all)"]
    for k, (n, t, x, fl) in rows:
        lines.append(f"{k[0]:10s} {k[1]:6d} {k[2]:6d} {k[3]:6d} {k[4]:2d} {k[5]:2d} {k[6]:3d} {k[7]:3d} {k[8]:6s} {n:5d} {t / n:9.1f} {x / n:8.1f} "
                     f"{(t + x) / 1e3:9.3f} {fl * n / (t + x) / 1e6:9.1f}")
    out = "\n".join(lines)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "gemm_in_step.txt"), "w").write(out + "\n")
    print(out)


if __name__ == "__main__":
    main()
