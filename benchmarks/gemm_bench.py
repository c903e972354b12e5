#!/usr/bin/env python
"""tcgen05 GEMM (``ops/gemm.py``) vs cuBLAS (torch.matmul) on the BERT-large shapes, forward AND backward.

    python benchmarks/gemm_bench.py [--quick]

For every (M, N, K) the three GEMMs of a linear layer are timed: forward ``X.W^T`` (+bias / +GELU), dgrad ``dY.W`` and
wgrad ``dY^T.X`` (the same FLOPs each).  CUDA events, median of 20, L2 flushed between iterations (256 MB memset).
The cuBLAS arm is what torch dispatches for the same math (``F.linear`` / ``@`` on the same, untransposed tensors).
Writes gpurun_out/gemm_bench.json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

SHAPES = [  # (M = tokens, N = out features, K = in features)
    (4096, 1024, 1024), (4096, 4096, 1024), (4096, 1024, 4096), (4096, 30528, 1024), (8192, 4096, 4096),
    (8192, 1024, 1024), (8192, 4096, 1024), (8192, 1024, 4096), (8192, 3072, 1024),
]


def timeit(fn, flush, iters=20, warm=5):
    for _ in range(warm):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3          # median, us


def main():
    from spark_deep_learning_b200.ops import gemm
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    shapes = SHAPES[:3] if "--quick" in sys.argv else SHAPES
    for m, n, k in shapes:
        x = torch.randn(m, k, device=dev, dtype=torch.bfloat16)
        w = torch.randn(n, k, device=dev, dtype=torch.bfloat16) * 0.05
        b = torch.randn(n, device=dev, dtype=torch.float32)
        wb = b.to(torch.bfloat16)
        dy = torch.randn(m, n, device=dev, dtype=torch.bfloat16)
        pre = torch.randn(m, k, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * m * n * k
        cases = [
            ("fwd+bias", lambda: gemm.gemm_bf16(x, w, b, None), lambda: F.linear(x, w, wb)),
            ("fwd+bias+gelu", lambda: gemm.gemm_bf16(x, w, b, "gelu", save_pre_act=True), lambda: F.gelu(F.linear(x, w, wb))),
            ("dgrad", lambda: gemm.dgrad(dy, w), lambda: dy @ w),
            ("dgrad*dgelu", lambda: gemm.dgrad(dy, w, pre, "gelu"), None),
            ("wgrad", lambda: gemm.wgrad(dy, x), lambda: dy.t() @ x),
        ]
        for name, ours, lib in cases:
            if n > 8192 and name in ("fwd+bias+gelu", "dgrad*dgelu"):
                continue
            t_ours = timeit(ours, flush)
            t_lib = timeit(lib, flush) if lib is not None else float("nan")
            rows.append({"M": m, "N": n, "K": k, "op": name, "ours_us": round(t_ours, 2), "lib_us": round(t_lib, 2),
                         "ours_tflops": round(fl / t_ours / 1e6, 1), "lib_tflops": round(fl / t_lib / 1e6, 1),
                         "ours_over_lib": round(t_lib / t_ours, 3)})
            print(rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_bench.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
