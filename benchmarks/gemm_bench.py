#!/usr/bin/env python
"""tcgen05 GEMM vs cuBLAS (torch) on the BERT-large forward shapes and one large square-ish shape.

    python benchmarks/gemm_bench.py                     # default kernel
    SPARKDL_GEMM_EPI16=1 python benchmarks/gemm_bench.py # experimental 16-epilogue-warp variant

CUDA-event timing, L2 flushed between iterations (256 MB memset), numerics checked against an fp32 reference first.
Writes gpurun_out/gemm_bench[_epi16].json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

SHAPES = [  # (M, N, K, act)
    (8192, 3072, 1024, None), (8192, 1024, 1024, None), (8192, 4096, 1024, "gelu"), (8192, 1024, 4096, None),
    (8192, 30528, 1024, None), (8192, 4096, 4096, None), (8192, 4096, 4096, "gelu"),
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
    for m, n, k, act in SHAPES:
        x = torch.randn(m, k, device=dev, dtype=torch.bfloat16)
        w = torch.randn(n, k, device=dev, dtype=torch.bfloat16) * 0.05
        b = torch.randn(n, device=dev, dtype=torch.float32)
        ref = F.linear(x[:256].float(), w.float(), b)
        ref = F.gelu(ref) if act else ref
        out, _ = gemm.gemm_bf16(x, w, b, act, save_pre_act=bool(act))
        err = (out[:256].float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
        t_ours = timeit(lambda: gemm.gemm_bf16(x, w, b, act, save_pre_act=bool(act)), flush)
        wb = b.to(torch.bfloat16)
        t_lib = timeit(lambda: F.gelu(F.linear(x, w, wb)) if act else F.linear(x, w, wb), flush)
        fl = 2.0 * m * n * k
        rows.append({"M": m, "N": n, "K": k, "act": act or "none", "rel_err": err, "ours_us": t_ours, "lib_us": t_lib,
                     "ours_tflops": fl / t_ours / 1e6, "lib_tflops": fl / t_lib / 1e6})
        print(rows[-1], flush=True)
    tag = "_epi16" if os.environ.get("SPARKDL_GEMM_EPI16") == "1" else ""
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_bench%s.json" % tag), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
