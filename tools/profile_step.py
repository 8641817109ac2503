"""Per-kernel time breakdown of one training step (CUPTI via torch.profiler) -> gpurun_out/step_profile.json/.txt.

    python tools/profile_step.py [--batch 4] [--res 1024]
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import Cfg, synth_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--mode", default="v_prediction")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-profiler", action="store_true")
    ap.add_argument("--eager", action="store_true")
    args = ap.parse_args()
    from aozora_sdxl_training_b200.optimizers import RavenAdamW
    from aozora_sdxl_training_b200.trainer import SDXLTrainStep
    from aozora_sdxl_training_b200.unet import UNet2DConditionModel, init_weights_fast_, sdxl_config
    dev = torch.device("cuda", 0)
    with torch.device(dev):
        unet = UNet2DConditionModel(sdxl_config()).to(torch.bfloat16)
    init_weights_fast_(unet)
    cfg = type("C", (Cfg,), dict(BATCH_SIZE=args.batch, PREDICTION_TYPE=args.mode, TIMESTEP_ALLOCATION=None))
    opt = RavenAdamW([{"params": list(unet.parameters()), "lr_scale": 1.0}], lr=8e-7, momentum_dtype=torch.bfloat16, **Cfg.RAVEN)
    step = SDXLTrainStep(unet, opt, cfg, device=dev, use_cuda_graph=not args.eager)
    batch = synth_batch(args.batch, args.res, 1, device=dev)
    for _ in range(args.warmup):
        step.step(batch)
    torch.cuda.synchronize()
    if args.no_profiler:
        # cudaProfilerStart/Stop bracket exactly the measured step(s): `ncu --profile-from-start off` lists only those launches
        torch.cuda.profiler.start()
        for _ in range(args.steps):
            step.step(batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    import time
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step.step(batch)
    e1.record()
    host_issue = (time.perf_counter() - t0) / 3
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / 3
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step.step(batch)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name
            agg[name][0] += 1
            agg[name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    rows = sorted(((n, c, t) for n, (c, t) in agg.items()), key=lambda r: -r[2])
    total = sum(r[2] for r in rows)
    out = dict(batch=args.batch, res=args.res, steps=args.steps, step_ms_events=wall, host_issue_ms=host_issue * 1e3,
               kernel_time_ms=total / 1e3 / args.steps,
               kernels=[dict(name=n[:120], launches=c // args.steps, ms=t / 1e3 / args.steps, share=t / total) for n, c, t in rows])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "step_profile.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "step_profile.txt"), "w") as f:
        f.write(f"step {wall:.2f} ms (events), host issue {host_issue * 1e3:.2f} ms, sum of kernel time {total / 1e3 / args.steps:.2f} ms\n")
        for n, c, t in rows[:40]:
            f.write(f"{t / 1e3 / args.steps:9.3f} ms {100 * t / total:5.1f}% x{c // args.steps:5d}  {n[:110]}\n")
    print(open(os.path.join(ROOT, "gpurun_out", "step_profile.txt")).read())


if __name__ == "__main__":
    main()
