"""Fused softmax cross-entropy over bf16 logits (``csrc/cuda/xent.cu``): one pass forward (per-row log-sum-exp), one
pass backward writing ``d(logits)`` in bf16 - instead of the fp32 copy + log_softmax + nll_loss chain (nine passes over
the ``[tokens, vocab]`` matrix).  ``cross_entropy(logits, labels)`` == ``F.cross_entropy(logits.float(), labels)``
(mean over all rows; no ignore_index rows in the BERT benchmark, negative labels are skipped)."""
import ctypes

import torch

from spark_deep_learning_b200.ops import comm as _comm
from spark_deep_learning_b200.utils import launch_counter

_sig = False


def _lib():
    global _sig
    lib = _comm.load()
    if not _sig:
        vp, i32, i64, u64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_float
        lib.sdlc_xent_fwd.restype = i32
        lib.sdlc_xent_fwd.argtypes = [vp, vp, vp, vp, i32, i32, i64, u64]
        lib.sdlc_xent_bwd.restype = i32
        lib.sdlc_xent_bwd.argtypes = [vp, vp, vp, vp, f32, vp, i32, i32, i64, u64]
        _sig = True
    return lib


def supported(logits, labels):
    """True when the fused sm_100a kernel can take these arguments (device, dtype, layout, alignment)."""
    return (logits.is_cuda and logits.dtype == torch.bfloat16 and logits.dim() == 2 and logits.stride(1) == 1
            and logits.shape[1] % 8 == 0 and logits.stride(0) % 8 == 0 and labels.dtype == torch.int64
            and labels.is_contiguous() and labels.numel() == logits.shape[0])


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class _XEnt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        t, v = logits.shape
        lse = torch.empty(t, dtype=torch.float32, device=logits.device)
        loss = torch.empty(t, dtype=torch.float32, device=logits.device)
        rc = _lib().sdlc_xent_fwd(_p(logits), _p(labels), _p(lse), _p(loss), t, v, logits.stride(0),
                                  torch.cuda.current_stream(logits.device).cuda_stream)
        if rc != 0:
            raise RuntimeError("sdlc_xent_fwd failed (%d)" % rc)
        launch_counter.add(1)
        ctx.save_for_backward(logits, labels, lse)
        return loss.mean()

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse = ctx.saved_tensors
        t, v = logits.shape
        d = torch.empty_like(logits)
        g = g.reshape(1).float().contiguous()
        rc = _lib().sdlc_xent_bwd(_p(logits), _p(labels), _p(lse), _p(g), 1.0 / t, _p(d), t, v, logits.stride(0),
                                  torch.cuda.current_stream(logits.device).cuda_stream)
        if rc != 0:
            raise RuntimeError("sdlc_xent_bwd failed (%d)" % rc)
        launch_counter.add(1)
        return d, None


def cross_entropy(logits, labels):
    """Mean softmax cross-entropy of bf16 ``logits`` [T, V] against int64 ``labels`` [T] (fp32 math, fp32 result)."""
    if not supported(logits, labels):
        return torch.nn.functional.cross_entropy(logits.float(), labels)
    return _XEnt.apply(logits, labels)
