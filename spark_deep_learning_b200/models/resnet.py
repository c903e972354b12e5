"""ResNet-50 (v1.5: stride on the 3x3 conv) for the flagship data-parallel benchmark
(BASELINE.json config 2: "ResNet-50 torch DistributedOptimizer via HorovodRunner(np=-8)").

Layout is NHWC (``channels_last``) bf16 end to end.  Convolutions are plain library calls
(cuDNN); everything between two convolutions - batch-norm statistics, normalise, residual add,
ReLU - runs in the fused sm_100a kernels of :mod:`spark_deep_learning_b200.ops.norm` when
``fused=True`` (one read + one write of the activation instead of five).
"""
import torch
from torch import nn


import os

# Forked block outputs (see Bottleneck.forward): opt-in.  Measured on B200 (benchmarks/gpu_batch.sh forklab, round 2):
# the in-kernel add removes autograd's add pass (1.3 ms / step) but the extra input stream costs the two backward
# kernels of the largest BatchNorm layers more than that (21.5 ms / step at best vs 21.1 ms without).
_FORK = os.environ.get("SPARKDL_BN_FORK", "0") == "1"


class _ConvBN(nn.Module):
    """conv -> batch-norm (-> +residual) (-> ReLU) with an optional fused epilogue."""

    def __init__(self, cin, cout, k, stride=1, relu=True, fused=False, zero_init=False):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        if zero_init:
            nn.init.zeros_(self.bn.weight)
        self.relu = relu
        self.fused = fused
        nn.init.kaiming_normal_(self.conv.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x, residual=None, fork=False):
        y = self.conv(x)
        if self.fused and y.is_cuda and self.training:
            from spark_deep_learning_b200.ops import norm
            # num_batches_tracked of all layers is advanced by ResNet.forward in one batched op
            return norm.batch_norm_act(y, self.bn, residual, self.relu, fork=fork, count_batches=False)
        y = self.bn(y)
        if residual is not None:
            y = y + residual
        y = torch.relu_(y) if self.relu else y
        return (y, y) if fork else y


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, width, stride, fused):
        super().__init__()
        cout = width * self.expansion
        self.a = _ConvBN(cin, width, 1, fused=fused)
        self.b = _ConvBN(width, width, 3, stride=stride, fused=fused)
        self.c = _ConvBN(width, cout, 1, relu=True, fused=fused, zero_init=True)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = _ConvBN(cin, cout, 1, stride=stride, relu=False, fused=fused)

    def forward(self, x):
        """``x`` is a tensor or a pair of tensors on one storage (the previous block's forked output): the first feeds
        the convolution branch, the second the identity / downsample branch, so that the two gradients reach the
        previous block's BatchNorm backward separately and are added inside its kernels (no autograd add pass).
        Returns such a pair."""
        xa, xi = x if isinstance(x, tuple) else (x, x)
        idt = xi if self.down is None else self.down(xi)
        return self.c(self.b(self.a(xa)), residual=idt, fork=_FORK)


class ResNet(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), num_classes=1000, fused=False):
        super().__init__()
        self.stem = _ConvBN(3, 64, 7, stride=2, fused=fused)
        self.pool = nn.MaxPool2d(3, stride=2, padding=1)
        self.fused = fused
        blocks, cin = [], 64
        for i, n in enumerate(layers):
            width = 64 * 2 ** i
            for j in range(n):
                blocks.append(Bottleneck(cin, width, (2 if i > 0 else 1) if j == 0 else 1, fused))
                cin = width * Bottleneck.expansion
        self.blocks = nn.Sequential(*blocks)
        self.fc = nn.Linear(cin, num_classes)

    def _stem(self, x):
        """7x7 / stride-2 stem.  On the fused path the input goes through one space-to-depth kernel (which also does
        the bf16 / channels-last cast) and the convolution runs as a dense 4x4 / stride-1 convolution over 16
        channels - the same arithmetic (the 7x7 weights are re-indexed, zero taps added), 3.1x faster in cuDNN than
        the 3-channel original (benchmarks/stem_bench.py)."""
        if self.fused and self.training:
            from spark_deep_learning_b200.ops import norm, pool
            if pool.stem_s2d_supported(x) and self.stem.conv.weight.dtype == torch.bfloat16:
                w = pool.stem_s2d_weight(self.stem.conv.weight).contiguous(memory_format=torch.channels_last)
                y = torch.nn.functional.conv2d(pool.stem_s2d(x), w)
                return norm.batch_norm_act(y, self.stem.bn, None, True, count_batches=False)
        if x.is_cuda and x.dtype != self.stem.conv.weight.dtype:
            x = x.to(dtype=self.stem.conv.weight.dtype, memory_format=torch.channels_last)
        return self.stem(x)

    def forward(self, x):
        x = self._stem(x)
        if self.fused and x.is_cuda:
            from spark_deep_learning_b200.ops import pool
            x = pool.max_pool_3x3_s2(x)        # 1-byte argmax, gather backward (falls back when unsupported)
        else:
            x = self.pool(x)
        x = self.blocks(x)
        if isinstance(x, tuple):
            x = x[0]
        if self.fused and self.training and x.is_cuda:
            nbt = [m.num_batches_tracked for m in self.modules()
                   if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None]
            if nbt:
                torch._foreach_add_(nbt, 1)          # one launch for the 53 counters
        x = torch.flatten(torch.nn.functional.adaptive_avg_pool2d(x, 1), 1)
        return self.fc(x)


def to_bf16_except_norm(model):
    """bf16 convolution / linear weights with fp32 BatchNorm parameters and statistics.

    With ``hvd.DistributedOptimizer(fused=True)`` the fp32 master copy of every bf16 weight lives in the sharded
    optimizer state, so no autocast (and none of its per-step weight casts) is needed: feed bf16 inputs."""
    model.to(torch.bfloat16)
    for m in model.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            m.float()
    return model


def resnet50(num_classes=1000, fused=False, bf16_weights=False):
    """ResNet-50 v1.5; ``fused`` selects the fused BN/ReLU/pool kernels, ``bf16_weights`` bf16 conv/fc weights with fp32 BN."""
    m = ResNet((3, 4, 6, 3), num_classes, fused)
    return to_bf16_except_norm(m) if bf16_weights else m


def resnet18ish(num_classes=10, fused=False, bf16_weights=False):
    """Tiny bottleneck net for smoke tests."""
    m = ResNet((1, 1, 1, 1), num_classes, fused)
    return to_bf16_except_norm(m) if bf16_weights else m
