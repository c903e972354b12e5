#!/usr/bin/env python
"""ResNet stem (7x7 / stride 2, 3 -> 64) as a 4x4 / stride 1 convolution on the space-to-depth input (2x2 pixel blocks,
channels padded 12 -> 16): numerics vs the direct convolution and cuDNN time of both (fprop, wgrad)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from benchmarks.gemm_bench import timeit  # noqa: E402


def s2d_input(x, cpad=4):
    """[N,3,H,W] -> [N,4*cpad,H/2+3,W/2+3] (zero padded: 2 left/top, 1 right/bottom), channel = (rh, rw, c)"""
    n, c, h, w = x.shape
    xp = F.pad(x, (0, 0, 0, 0, 0, cpad - c))
    x2 = xp.view(n, cpad, h // 2, 2, w // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * cpad, h // 2, w // 2)
    return F.pad(x2, (2, 1, 2, 1))


def s2d_weight(w, cpad=4):
    """[O,3,7,7] -> [O,4*cpad,4,4]: tap i = 2a + r + 3 with a in -2..1 (kernel row a+2), r in 0..1"""
    o, c = w.shape[:2]
    w8 = F.pad(w, (1, 0, 1, 0, 0, cpad - c))            # [O,cpad,8,8]: index i+1 = 2(a+2) + r
    return w8.view(o, cpad, 4, 2, 4, 2).permute(0, 3, 5, 1, 2, 4).reshape(o, 4 * cpad, 4, 4)


def main():
    torch.backends.cudnn.benchmark = True
    n = 256
    x = torch.randn(n, 3, 224, 224, device="cuda")
    w = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
    ref = F.conv2d(x[:8], w, stride=2, padding=3)
    alt = F.conv2d(s2d_input(x[:8]), s2d_weight(w))
    print("max abs diff fp32:", float((ref - alt).abs().max()), "shape", tuple(alt.shape))
    xb = x.bfloat16().contiguous(memory_format=torch.channels_last)
    wb = w.bfloat16().contiguous(memory_format=torch.channels_last)
    x2 = s2d_input(x).bfloat16().contiguous(memory_format=torch.channels_last)
    w2 = s2d_weight(w).bfloat16().contiguous(memory_format=torch.channels_last)
    dy = torch.randn(n, 64, 112, 112, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    print("direct 7x7 s2 : fprop %.1f us" % timeit(lambda i: F.conv2d(xb, wb, stride=2, padding=3), 1, rounds=3))
    print("s2d   4x4 s1 : fprop %.1f us" % timeit(lambda i: F.conv2d(x2, w2), 1, rounds=3))
    bw = torch.ops.aten.convolution_backward
    print("direct wgrad %.1f us" % timeit(lambda i: bw(dy, xb, wb, None, [2, 2], [3, 3], [1, 1], False, [0, 0], 1, [False, True, False]), 1, rounds=3))
    print("s2d    wgrad %.1f us" % timeit(lambda i: bw(dy, x2, w2, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [False, True, False]), 1, rounds=3))
    xin = torch.randn(n, 3, 224, 224, device="cuda")
    print("input transform (fp32 NCHW -> bf16 s2d NHWC): %.1f us" % timeit(lambda i: s2d_input(xin).to(dtype=torch.bfloat16, memory_format=torch.channels_last), 1, rounds=3))
    print("input cast only (fp32 NCHW -> bf16 NHWC): %.1f us" % timeit(lambda i: xin.to(dtype=torch.bfloat16, memory_format=torch.channels_last), 1, rounds=3))
    for cpad in (8,):
        x2 = s2d_input(x, cpad).bfloat16().contiguous(memory_format=torch.channels_last)
        w2 = s2d_weight(w, cpad).bfloat16().contiguous(memory_format=torch.channels_last)
        print("s2d cpad=%d (C=%d): fprop %.1f us  wgrad %.1f us" % (cpad, 4 * cpad, timeit(lambda i: F.conv2d(x2, w2), 1, rounds=3),
              timeit(lambda i: bw(dy, x2, w2, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [False, True, False]), 1, rounds=3)))


if __name__ == "__main__":
    main()
