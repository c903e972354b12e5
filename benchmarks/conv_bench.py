#!/usr/bin/env python
"""Is the tcgen05 GEMM (+ BN-statistics epilogue) competitive with cuDNN on ResNet-50's 1x1 convolutions (NHWC bf16,
batch 256)?  Also: the 7x7 stem with the input padded to 4 / 8 channels.  CUDA-graph timed like gemm_bench.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from benchmarks.gemm_bench import timeit  # noqa: E402


def main():
    from spark_deep_learning_b200.ops import gemm, norm
    torch.backends.cudnn.benchmark = True
    n = 256
    shapes = [(64, 64, 56), (64, 256, 56), (256, 64, 56), (256, 128, 56), (128, 512, 28), (512, 128, 28), (512, 256, 28),
              (256, 1024, 14), (1024, 256, 14), (1024, 512, 14), (512, 2048, 7), (2048, 512, 7)]
    for cin, cout, hw in shapes:
        nsets = 3
        xs = [torch.randn(n, cin, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(nsets)]
        ws = [(torch.randn(cout, cin, 1, 1, device="cuda") * 0.05).bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(nsets)]
        dys = [torch.randn(n, cout, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(nsets)]
        m = n * hw * hw
        x2 = [x.permute(0, 2, 3, 1).reshape(m, cin) for x in xs]
        w2 = [w.reshape(cout, cin) for w in ws]
        dy2 = [d.permute(0, 2, 3, 1).reshape(m, cout) for d in dys]
        bn = torch.nn.BatchNorm2d(cout).cuda()
        t_cudnn_f = timeit(lambda i: F.conv2d(xs[i], ws[i]), nsets, rounds=2)

        t_ours_f = timeit(lambda i: gemm.gemm_ex(x2[i], w2[i], m, cout, cin, epi=gemm.EPI_ACC_BF16), nsets, rounds=2)
        t_ours_fs = timeit(lambda i: gemm.conv1x1_stats(x2[i], w2[i]), nsets, rounds=2)
        t_cudnn_d = timeit(lambda i: torch.ops.aten.convolution_backward(dys[i], xs[i], ws[i], None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False]), nsets, rounds=2)
        t_ours_d = timeit(lambda i: gemm.dgrad(dy2[i], w2[i]), nsets, rounds=2)
        t_cudnn_w = timeit(lambda i: torch.ops.aten.convolution_backward(dys[i], xs[i], ws[i], None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [False, True, False]), nsets, rounds=2)
        t_ours_w = timeit(lambda i: gemm.wgrad(dy2[i], x2[i]), nsets, rounds=2)
        print("1x1 %4d->%4d @%3d  fwd cudnn %6.1f ours %6.1f ours+stats %6.1f | dgrad cudnn %6.1f ours %6.1f | wgrad cudnn %6.1f ours %6.1f  (us)"
              % (cin, cout, hw, t_cudnn_f, t_ours_f, t_ours_fs, t_cudnn_d, t_ours_d, t_cudnn_w, t_ours_w), flush=True)
        del xs, ws, dys, x2, w2, dy2
        torch.cuda.empty_cache()
    # stem: 7x7 / 2, 3 -> 64 channels, input padded to 4 / 8 channels
    for cpad in (3, 4, 8):
        x = torch.randn(n, cpad, 224, 224, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        w = (torch.randn(64, cpad, 7, 7, device="cuda") * 0.05).bfloat16().contiguous(memory_format=torch.channels_last)
        dy = torch.randn(n, 64, 112, 112, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        tf = timeit(lambda i: F.conv2d(x, w, stride=2, padding=3), 1, rounds=3)
        tw = timeit(lambda i: torch.ops.aten.convolution_backward(dy, x, w, None, [2, 2], [3, 3], [1, 1], False, [0, 0], 1, [False, True, False]), 1, rounds=3)
        print("stem 7x7 s2 cin=%d: fprop %.1f us  wgrad %.1f us" % (cpad, tf, tw), flush=True)


if __name__ == "__main__":
    main()
