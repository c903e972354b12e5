"""3x3 / stride-2 / pad-1 max pooling for channels-last bf16 activations (``csrc/cuda/pool.cu``): 1-byte argmax,
gather-style backward.  Falls back to ``F.max_pool2d`` for anything else."""
import ctypes

import torch

from spark_deep_learning_b200.ops import comm as _comm
from spark_deep_learning_b200.utils import launch_counter

_sig = False


def _lib():
    global _sig
    lib = _comm.load()
    if not _sig:
        vp, i32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_ulonglong
        lib.sdlc_maxpool3s2_fwd.restype = i32
        lib.sdlc_maxpool3s2_fwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, u64]
        lib.sdlc_maxpool3s2_bwd.restype = i32
        lib.sdlc_maxpool3s2_bwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, u64]
        _sig = True
    return lib


def supported(x):
    """True when the fused sm_100a kernel can take these arguments (device, dtype, layout, alignment)."""
    return (x.is_cuda and x.dim() == 4 and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last))


class _MaxPool3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        n, c, h, w = x.shape
        oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        y = torch.empty((n, c, oh, ow), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device)
        rc = _lib().sdlc_maxpool3s2_fwd(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()),
                                        ctypes.c_void_p(idx.data_ptr()), n, h, w, c,
                                        torch.cuda.current_stream(x.device).cuda_stream)
        if rc != 0:
            raise RuntimeError("sdlc_maxpool3s2_fwd failed (%d)" % rc)
        launch_counter.add(1)
        ctx.save_for_backward(idx)
        ctx.shape = (n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        n, c, h, w = ctx.shape
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((n, c, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        rc = _lib().sdlc_maxpool3s2_bwd(ctypes.c_void_p(dy.data_ptr()), ctypes.c_void_p(idx.data_ptr()),
                                        ctypes.c_void_p(dx.data_ptr()), n, h, w, c,
                                        torch.cuda.current_stream(dy.device).cuda_stream)
        if rc != 0:
            raise RuntimeError("sdlc_maxpool3s2_bwd failed (%d)" % rc)
        launch_counter.add(1)
        return dx


def max_pool_3x3_s2(x):
    """3x3 / stride-2 / pad-1 max-pool for NHWC tensors (1-byte argmax, gather backward); falls back to torch when unsupported."""
    if not supported(x):
        return torch.nn.functional.max_pool2d(x, 3, stride=2, padding=1)
    return _MaxPool3s2.apply(x)
