"""Raven update over the REAL SDXL parameter table (1680 tensors, 2,567,463,684 parameters, bf16 p / g / m / v = 14 B per parameter):
the bulk-copy streaming kernel against the register-streaming kernel, CUDA-event timed.   python tools/raven_table_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from aozora_sdxl_training_b200 import _lib  # noqa: E402
from aozora_sdxl_training_b200.optimizers import RavenAdamW  # noqa: E402
from aozora_sdxl_training_b200.unet import UNet2DConditionModel, sdxl_config  # noqa: E402


def main():
    with torch.device("meta"):
        numels = [p.numel() for p in UNet2DConditionModel(sdxl_config()).parameters()]
    ps = [torch.nn.Parameter(torch.zeros(n, device="cuda", dtype=torch.bfloat16)) for n in numels]
    for p in ps:
        p.grad = torch.full_like(p, 1e-3)
    opt = RavenAdamW(ps, lr=8e-7, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, debias_strength=0.3)

    def timeit(n=5):
        for _ in range(2):
            opt.step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            opt.step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    tot = sum(numels)
    for mode in (1, 0, 1, 0):
        _lib.call("aoz_raven_set_bulk", mode)
        ms = timeit()
        print(("bulk copies  " if mode else "register path"), f"{ms:.3f} ms  {14.0 * tot / ms / 1e6:.1f} GB/s", flush=True)
    _lib.call("aoz_raven_set_bulk", 1)


if __name__ == "__main__":
    main()
