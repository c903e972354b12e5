#!/usr/bin/env python
"""Epilogue experiments on the tcgen05 GEMM: bias kind (none / fp32 / bf16) x activation x operand residency
(warm = same operands every launch, L2 resident; cold = a ring of operand sets larger than L2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from spark_deep_learning_b200.ops import gemm  # noqa: E402

SHAPES = [(4096, 4096, 1024), (4096, 1024, 1024), (4096, 1024, 4096)]


def timeit(fn, n=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def main():
    for m, n, k in SHAPES:
        sets = max(2, int(400e6 // ((m * k + n * k + m * n) * 2)) + 1)
        xs = [torch.randn(m, k, device="cuda").bfloat16() for _ in range(sets)]
        ws = [(torch.randn(n, k, device="cuda") * 0.05).bfloat16() for _ in range(sets)]
        outs = [torch.empty(m, n, device="cuda", dtype=torch.bfloat16) for _ in range(sets)]
        b32 = torch.randn(n, device="cuda")
        b16 = b32.bfloat16()
        for res in ("warm", "cold"):
            for bias, bname in ((None, "nobias"), (b32, "bias32"), (b16, "bias16")):
                for act in (None, "gelu"):
                    def run(i):
                        j = (i % sets) if res == "cold" else 0
                        gemm.gemm_ex(xs[j], ws[j], m, n, k, epi=gemm.EPI_LINEAR, out=outs[j], bias=bias, act=act)
                    us = timeit(run)
                    print("M%d N%d K%d %s %s %s: %.1f us %.0f TFLOP/s" % (m, n, k, res, bname, act, us, 2.0 * m * n * k / us / 1e6), flush=True)
            def lib(i):
                j = (i % sets) if res == "cold" else 0
                torch.mm(xs[j], ws[j].t(), out=outs[j])
            us = timeit(lib)
            print("M%d N%d K%d %s cublas: %.1f us %.0f TFLOP/s" % (m, n, k, res, us, 2.0 * m * n * k / us / 1e6), flush=True)
            def libb(i):
                j = (i % sets) if res == "cold" else 0
                torch.addmm(b16, xs[j], ws[j].t(), out=outs[j])
            us = timeit(libb)
            print("M%d N%d K%d %s cublas+bias: %.1f us %.0f TFLOP/s" % (m, n, k, res, us, 2.0 * m * n * k / us / 1e6), flush=True)


if __name__ == "__main__":
    main()
