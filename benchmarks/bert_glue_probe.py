#!/usr/bin/env python
"""One-GPU probe of the PyTorch glue left in the BERT-large layer (NOTES_NEXT_ROUND.md item 5).

Answers, with CUDA-event timings at BERT-large shapes (b=64, s=128, h=1024, 16 heads):
 A. packed-QKV attention (one [h,3h] GEMM, q/k/v strided views) vs three separate projections: forward+backward time,
    and whether dq/dk/dv come back from the attention backward as views of ONE packed tensor (then `stack` is free);
 B. whether `a.transpose(1, 2).reshape(b, s, h)` after SDPA copies (output layout of the library kernel);
 C. bias-gradient column sums: `dy.sum(0)` vs `ones @ dy` (GEMV through the tensor cores) at N=1024 / 3072 / 4096.
Writes gpurun_out/bert_glue_probe.json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


CPU_DRY_RUN = os.environ.get("PROBE_CPU") == "1"     # syntax / shape check without a GPU


def timeit(fn, iters=20, warm=5):
    if CPU_DRY_RUN:
        fn()
        return 0.0
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3          # us


def main():
    dev = torch.device("cpu" if CPU_DRY_RUN else "cuda")
    b, s, h, nh = (2, 16, 64, 4) if CPU_DRY_RUN else (64, 128, 1024, 16)
    d = h // nh
    out = {}
    x = torch.randn(b, s, h, device=dev, dtype=torch.bfloat16, requires_grad=True)
    wqkv = torch.randn(3 * h, h, device=dev, dtype=torch.bfloat16, requires_grad=True) * 0.02
    wqkv = wqkv.detach().requires_grad_()
    ws = [wqkv.detach()[i * h:(i + 1) * h].clone().requires_grad_() for i in range(3)]
    g = torch.randn(b, s, h, device=dev, dtype=torch.bfloat16)

    def packed():
        q, k, v = F.linear(x, wqkv).view(b, s, 3, nh, d).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v)
        a.transpose(1, 2).reshape(b, s, h).backward(g)
        x.grad = wqkv.grad = None

    def split():
        q, k, v = [F.linear(x, w).view(b, s, nh, d).transpose(1, 2) for w in ws]
        a = F.scaled_dot_product_attention(q, k, v)
        a.transpose(1, 2).reshape(b, s, h).backward(g)
        x.grad = None
        for w in ws:
            w.grad = None

    out["A_packed_us"] = timeit(packed)
    out["A_split_us"] = timeit(split)

    # layouts: does the attention backward hand back one packed tensor?  is the output bshd in memory?
    qkv = F.linear(x, wqkv).view(b, s, 3, nh, d).detach().requires_grad_()
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    q.retain_grad(), k.retain_grad(), v.retain_grad()
    a = F.scaled_dot_product_attention(q, k, v)
    out["B_attn_out_strides"] = list(a.stride())
    out["B_reshape_is_view"] = a.transpose(1, 2).reshape(b, s, h).data_ptr() == a.data_ptr() and \
        a.transpose(1, 2).is_contiguous()
    a.backward(torch.randn_like(a))
    out["A_dq_strides"] = list(q.grad.stride())
    base = q.grad.untyped_storage().data_ptr()
    out["A_dqkv_share_storage"] = base == k.grad.untyped_storage().data_ptr() == v.grad.untyped_storage().data_ptr()
    out["A_dq_contig_as_bshd"] = q.grad.transpose(1, 2).is_contiguous()

    for backend in ("CUDNN_ATTENTION", "FLASH_ATTENTION", "EFFICIENT_ATTENTION"):
        try:
            from torch.nn.attention import SDPBackend, sdpa_kernel
            with sdpa_kernel(getattr(SDPBackend, backend)):
                out["A_packed_us_" + backend] = timeit(packed)
        except Exception as e:  # pylint: disable=broad-except
            out["A_packed_us_" + backend] = "error: %s" % str(e)[:120]

    # C. bias gradients
    for n in (1024, 3072, 4096):
        dy = torch.randn(b * s, n, device=dev, dtype=torch.bfloat16)
        ones = torch.ones(1, b * s, device=dev, dtype=torch.bfloat16)
        out["C_sum0_us_N%d" % n] = timeit(lambda: dy.sum(0))
        out["C_ones_gemm_us_N%d" % n] = timeit(lambda: ones @ dy)
        out["C_bytes_N%d" % n] = dy.numel() * 2
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bert_glue_probe.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
