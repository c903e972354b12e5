#!/usr/bin/env python
"""Pipeline experiments on the tcgen05 GEMM: kernel time for a few shapes, 20 back-to-back launches between two CUDA
events (the CPU runs ahead, so this is device time).  Combine with SPARKDL_GEMM_DBG=1|2|3|4 (no TMA / no MMA / neither /
no epilogue stores) to see which side of the pipeline sets the pace.

    python benchmarks/gemm_lab.py [ncu]      # "ncu": a handful of launches only, for a profiler capture
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from spark_deep_learning_b200.ops import gemm  # noqa: E402

SHAPES = [(4096, 1024, 4096), (4096, 4096, 1024), (8192, 4096, 4096)]


def main():
    ncu = "ncu" in sys.argv
    dbg = os.environ.get("SPARKDL_GEMM_DBG", "0")
    for m, n, k in SHAPES:
        x = torch.randn(m, k, device="cuda").bfloat16()
        w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        for bn in (256, 128):
            def run():
                gemm.gemm_ex(x, w, m, n, k, epi=gemm.EPI_LINEAR, out=out, block_n=bn)
            for _ in range(2 if ncu else 5):
                run()
            if ncu:
                continue
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            print("dbg=%s M%d N%d K%d bn%d: %.1f us  %.0f TFLOP/s" % (dbg, m, n, k, bn, us, 2.0 * m * n * k / us / 1e6), flush=True)
        if not ncu:
            xx, ww = x, w
            for _ in range(5):
                xx @ ww.t()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                xx @ ww.t()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            print("cublas M%d N%d K%d: %.1f us  %.0f TFLOP/s" % (m, n, k, us, 2.0 * m * n * k / us / 1e6), flush=True)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
