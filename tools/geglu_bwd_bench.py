"""GEGLU backward + the projection's bias gradient: separate launches (geglu_bwd + colsum) against the fused kernel at several
row-chunk counts, CUDA-graph timed at the SDXL shapes.   python tools/geglu_bwd_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from aozora_sdxl_training_b200 import _lib, ops  # noqa: E402

BF = torch.bfloat16


def timeit_graph(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    for M, C in ((4096, 1280), (16384, 640)):
        half = 4 * C
        dy = torch.randn(M, half, device="cuda").to(BF)
        aux = torch.randn(M, 2 * half, device="cuda").to(BF)
        db = torch.empty(2 * half, device="cuda", dtype=BF)

        def separate():
            d = ops.geglu_bwd(dy, aux)
            ops.colsum(d, out=db)

        print(f"M={M} C={C}: separate {timeit_graph(separate):.1f} us", end="")
        for mul in (4, 8, 16, 32):
            _lib.call("aoz_geglu_colsum_set_blocks_per_sm", mul)
            print(f" | fused x{mul} {timeit_graph(lambda: ops.geglu_bwd(dy, aux, bias_grad=db)):.1f} us", end="")
        print(flush=True)
    _lib.call("aoz_geglu_colsum_set_blocks_per_sm", 4)


if __name__ == "__main__":
    main()
