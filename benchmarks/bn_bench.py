#!/usr/bin/env python
"""Per-kernel time and achieved HBM bandwidth of the fused BatchNorm / max-pool kernels on ResNet-50 shapes
(torch.profiler kernel times, inputs larger than L2).  python benchmarks/bn_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    from spark_deep_learning_b200.ops import norm, pool
    rows = []
    for (n, c, h, w, res) in [(256, 256, 56, 56, True), (256, 64, 56, 56, False), (256, 512, 28, 28, True),
                              (256, 1024, 14, 14, True), (256, 2048, 7, 7, True), (256, 64, 112, 112, False)]:
        x = torch.randn(n, c, h, w, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        r = torch.randn_like(x).requires_grad_(True) if res else None
        m = torch.nn.BatchNorm2d(c).cuda()
        g = torch.randn_like(x)
        for _ in range(3):
            y = norm.batch_norm_act(x, m, r, True)
            y.backward(g)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                y = norm.batch_norm_act(x, m, r, True)
                y.backward(g)
            torch.cuda.synchronize()
        t = x.numel() * 2 / 1e6   # MB per pass over the activation
        traffic = {"bn_stats": t, "bn_apply": (3 if res else 2) * t + t / 16, "bn_bwd_reduce": 2 * t + t / 16,
                   "bn_bwd_apply": (4 if res else 3) * t + t / 16}
        for e in prof.key_averages():
            for k, mb in traffic.items():
                if ("bn::%s_kernel" % k) in e.key:
                    us = e.device_time_total / e.count
                    rows.append("%-14s n%d c%d %dx%d res=%d  %7.1f us  %6.0f GB/s (%.0f%% of 6571)" % (
                        k, n, c, h, w, int(res), us, mb / us * 1e3, mb / us * 1e3 / 65.71))
        del x, r, g, y
        torch.cuda.empty_cache()
    x = torch.randn(256, 64, 112, 112, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = None
    for _ in range(3):
        y = pool.max_pool_3x3_s2(x)
        g = torch.randn_like(y) if g is None else g
        y.backward(g)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            y = pool.max_pool_3x3_s2(x)
            y.backward(g)
        torch.cuda.synchronize()
    tin, tout = x.numel() * 2 / 1e6, y.numel() * 2 / 1e6
    for e in prof.key_averages():
        if "maxpool_fwd" in e.key:
            us = e.device_time_total / e.count
            rows.append("maxpool_fwd   %7.1f us  %6.0f GB/s" % (us, (tin + tout + tout / 2) / us * 1e3))
        if "maxpool_bwd" in e.key:
            us = e.device_time_total / e.count
            rows.append("maxpool_bwd   %7.1f us  %6.0f GB/s" % (us, (tin + tout + tout / 2) / us * 1e3))
    print("\n".join(rows))


if __name__ == "__main__":
    main()
