#!/usr/bin/env python
"""tcgen05 GEMM (``ops/gemm.py``) vs cuBLAS (torch.matmul) on the BERT-large shapes, forward AND backward.

    python benchmarks/gemm_bench.py [--quick]

For every (M, N, K) the three GEMMs of a linear layer are timed: forward ``X.W^T`` (+bias / +GELU), dgrad ``dY.W`` and
wgrad ``dY^T.X`` (the same FLOPs each).  Timing: CUDA events around 4 rounds over a ring of independent operand sets
whose total size exceeds the 126 MB L2 ("inputs larger than L2"), captured into one CUDA graph (no host launch cost in the number); 5 replays,
median.  The cuBLAS arm is what torch dispatches for the same math (``F.linear`` / ``@`` on the
same, untransposed tensors) and is timed the same way.  Writes gpurun_out/gemm_bench.json.
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


def timeit(fn, nsets, reps=5, rounds=4):
    """fn(i) launches the op on operand set i; returns median microseconds per launch.  The launches are captured
    into one CUDA graph (what the training step does as well), so host-side launch cost is not part of the number."""
    for i in range(nsets):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _r in range(rounds):
                for i in range(nsets):
                    fn(i)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / (rounds * nsets))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    from spark_deep_learning_b200.ops import gemm
    dev = torch.device("cuda")
    rows = []
    shapes = SHAPES[:3] if "--quick" in sys.argv else SHAPES
    for m, n, k in shapes:
        per_set = 2 * (m * k + n * k + 2 * m * n)                    # x, w, dy/out (bf16)
        nsets = max(2, min(12, (160 << 20) // per_set + 1))           # ring > L2
        xs = [torch.randn(m, k, device=dev, dtype=torch.bfloat16) for _ in range(nsets)]
        ws = [torch.randn(n, k, device=dev, dtype=torch.bfloat16) * 0.05 for _ in range(nsets)]
        dys = [torch.randn(m, n, device=dev, dtype=torch.bfloat16) for _ in range(nsets)]
        pres = [torch.randn(m, k, device=dev, dtype=torch.bfloat16) for _ in range(min(nsets, 4))]
        b = torch.randn(n, device=dev, dtype=torch.float32)
        wb = b.to(torch.bfloat16)
        fl = 2.0 * m * n * k
        cases = [
            ("fwd+bias", lambda i: gemm.gemm_bf16(xs[i], ws[i], b, None), lambda i: F.linear(xs[i], ws[i], wb)),
            ("fwd+bias+gelu", lambda i: gemm.gemm_bf16(xs[i], ws[i], b, "gelu", save_pre_act=True),
             lambda i: F.gelu(F.linear(xs[i], ws[i], wb))),
            ("dgrad", lambda i: gemm.dgrad(dys[i], ws[i]), lambda i: dys[i] @ ws[i]),
            ("dgrad*dgelu", lambda i: gemm.dgrad(dys[i], ws[i], pres[i % len(pres)], "gelu"), None),
            ("wgrad", lambda i: gemm.wgrad(dys[i], xs[i]), lambda i: dys[i].t() @ xs[i]),
        ]
        for name, ours, lib in cases:
            if n > 8192 and name in ("fwd+bias+gelu", "dgrad*dgelu"):
                continue
            # ours, library, ours, library: the better of two interleaved measurements each, so that neither arm is the
            # one that pays for the clock ramp after the (idle) operand set-up of a new shape
            t_ours = timeit(ours, nsets)
            t_lib = timeit(lib, nsets) if lib is not None else float("nan")
            t_ours = min(t_ours, timeit(ours, nsets))
            if lib is not None:
                t_lib = min(t_lib, timeit(lib, nsets))
            rows.append({"M": m, "N": n, "K": k, "op": name, "ours_us": round(t_ours, 2), "lib_us": round(t_lib, 2),
                         "ours_tflops": round(fl / t_ours / 1e6, 1), "lib_tflops": round(fl / t_lib / 1e6, 1),
                         "ours_over_lib": round(t_lib / t_ours, 3), "operand_sets": nsets})
            print(rows[-1], flush=True)
        del xs, ws, dys, pres
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_bench.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
