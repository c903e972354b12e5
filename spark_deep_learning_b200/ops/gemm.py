"""bf16 linear layer on the tcgen05 tensor cores with a fused bias / GELU epilogue
(``csrc/cuda/gemm_tcgen05.cu``).

``linear_act(x, weight, bias, act)`` == ``act(F.linear(x, weight, bias))`` for bf16 operands.
Forward runs the hand-written TMA -> tcgen05.mma -> TMEM -> epilogue kernel; backward uses plain
library GEMMs (cuBLAS through torch.matmul) for dgrad / wgrad plus the analytic GELU derivative on
the saved pre-activation.
"""
import ctypes
import os

import torch

from spark_deep_learning_b200.ops import comm as _comm
from spark_deep_learning_b200.utils import launch_counter

ACT = {None: 0, "none": 0, "gelu_tanh": 1, "gelu": 2, "gelu_erf": 2}
_sig = False
# EXPERIMENTAL: 16-epilogue-warp kernel variant (gemm_tcgen05_epi16.cu), opt-in until it has been run on hardware
_EPI16 = os.environ.get("SPARKDL_GEMM_EPI16", "0") == "1"


def _lib():
    global _sig
    lib = _comm.load()
    if not _sig:
        vp, i32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_ulonglong
        lib.sdlc_gemm_bf16.restype = i32
        lib.sdlc_gemm_bf16.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, u64]
        lib.sdlc_gelu_bwd.restype = i32
        lib.sdlc_gelu_bwd.argtypes = [vp, vp, vp, ctypes.c_longlong, i32, u64]
        if _EPI16:
            lib.sdlc_gemm_bf16_epi16.restype = i32
            lib.sdlc_gemm_bf16_epi16.argtypes = lib.sdlc_gemm_bf16.argtypes
        _sig = True
    return lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def supported(x, weight):
    """True when the fused sm_100a kernel can take these arguments (device, dtype, layout, alignment)."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
            and x.shape[-1] % 8 == 0 and weight.shape[0] % 8 == 0 and weight.shape[1] == x.shape[-1])


def gemm_bf16(x2d, weight, bias=None, act=None, save_pre_act=False, block_n=None):
    """Raw kernel call: x2d [M,K] bf16, weight [N,K] bf16, bias fp32 [N] or None -> [M,N] bf16."""
    lib = _lib()
    x2d = x2d.contiguous()
    weight = weight.contiguous()
    m, k = x2d.shape
    n = weight.shape[0]
    out = torch.empty((m, n), dtype=torch.bfloat16, device=x2d.device)
    pre = torch.empty_like(out) if save_pre_act else None
    if bias is not None and bias.dtype != torch.float32:
        bias = bias.float()
    if block_n is None:
        # enough tiles to fill 148 SMs with the wide tile? otherwise use the narrow one
        block_n = 256 if ((m + 127) // 128) * ((n + 255) // 256) >= 148 else 128
    fn = lib.sdlc_gemm_bf16_epi16 if _EPI16 else lib.sdlc_gemm_bf16
    rc = fn(_p(x2d), _p(weight), _p(bias), _p(out), _p(pre), m, n, k, ACT[act], block_n, 0,
            torch.cuda.current_stream(x2d.device).cuda_stream)
    if rc != 0:
        raise RuntimeError("sdlc_gemm_bf16 failed (%d) for M=%d N=%d K=%d" % (rc, m, n, k))
    launch_counter.add(1)
    return out, pre


class _LinearAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        x2d = x.reshape(-1, x.shape[-1])
        need_pre = act not in (None, "none") and (x.requires_grad or weight.requires_grad)
        out, pre = gemm_bf16(x2d, weight, bias, act, save_pre_act=need_pre)
        ctx.save_for_backward(x2d, weight, pre)
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return out.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2d, weight, pre = ctx.saved_tensors
        dy2d = dy.reshape(-1, dy.shape[-1])
        if pre is not None:
            dy2d = dy2d.contiguous()
            g = torch.empty_like(pre)
            rc = _lib().sdlc_gelu_bwd(_p(dy2d), _p(pre), _p(g), pre.numel(), 1 if ctx.act == "gelu_tanh" else 2,
                                      torch.cuda.current_stream(pre.device).cuda_stream)
            if rc != 0:
                raise RuntimeError("sdlc_gelu_bwd failed")
            launch_counter.add(1)
            dy2d = g
        dx = dy2d @ weight                                   # dgrad: plain library GEMM
        dw = dy2d.t() @ x2d                                  # wgrad: plain library GEMM
        db = dy2d.sum(0, dtype=torch.float32) if ctx.has_bias else None
        return dx.view(ctx.x_shape), dw, db, None


def linear_act(x, weight, bias=None, act=None):
    """``act(x @ weight.T + bias)``; bf16 in / out, fp32 accumulate, tcgen05 forward."""
    if not supported(x, weight):
        y = torch.nn.functional.linear(x, weight, None if bias is None else bias.to(x.dtype))
        if act in ("gelu", "gelu_erf"):
            return torch.nn.functional.gelu(y)
        if act == "gelu_tanh":
            return torch.nn.functional.gelu(y, approximate="tanh")
        return y
    return _LinearAct.apply(x, weight, bias, act)
