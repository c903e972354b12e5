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
        i64 = ctypes.c_longlong
        lib.sdlc_stem_s2d.restype = i32
        lib.sdlc_stem_s2d.argtypes = [vp, vp, i32, i32, i32, i64, i64, i64, i64, i32, u64]
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


def stem_s2d_supported(x):
    """True when :func:`stem_s2d` can take ``x`` ([N,3,H,W] fp32 / bf16 on a GPU, even H and W)."""
    return (x.is_cuda and x.dim() == 4 and x.shape[1] == 3 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0
            and x.dtype in (torch.float32, torch.bfloat16))


def stem_s2d(x):
    """Space-to-depth view of a 3-channel image batch for the ResNet stem: ``[N,3,H,W]`` (fp32 or bf16, NCHW or
    channels-last) -> bf16 channels-last ``[N,16,H/2+3,W/2+3]`` with channel ``(rh*2+rw)*4 + c`` holding pixel
    ``(2h'+rh-4, 2w'+rw-4)`` (zero outside the image and for c = 3).  One streaming kernel; it also replaces the
    fp32 -> bf16 channels-last cast of the input.  No gradient (the network input needs none)."""
    n, _, h, w = x.shape
    out = torch.empty((n, 16, h // 2 + 3, w // 2 + 3), dtype=torch.bfloat16, device=x.device,
                      memory_format=torch.channels_last)
    sn, sc, sh, sw = x.stride()
    rc = _lib().sdlc_stem_s2d(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), n, h, w, sn, sc, sh, sw,
                              6 if x.dtype == torch.float32 else 5, torch.cuda.current_stream(x.device).cuda_stream)
    if rc != 0:
        raise RuntimeError("sdlc_stem_s2d failed (%d)" % rc)
    launch_counter.add(1)
    return out


def stem_s2d_weight(w):
    """``[O,3,7,7]`` stem weights -> the equivalent ``[O,16,4,4]`` kernel over the space-to-depth input (plain torch
    ops on a tiny tensor, so autograd carries the weight gradient back to the original layout)."""
    o = w.shape[0]
    w8 = torch.nn.functional.pad(w, (1, 0, 1, 0, 0, 1))               # [O,4,8,8]; tap i+1 = 2(a+2) + r
    return w8.view(o, 4, 4, 2, 4, 2).permute(0, 3, 5, 1, 2, 4).reshape(o, 16, 4, 4)
