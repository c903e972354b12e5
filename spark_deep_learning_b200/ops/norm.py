"""Fused BatchNorm (+ residual add) (+ ReLU) for channels-last activations (``csrc/cuda/bn_act.cu``).

``batch_norm_act(x, bn_module, residual, relu)`` replaces ``relu(bn(x) + residual)`` in training
mode: two streaming passes in forward, two in backward, bf16 in/out, fp32 statistics, a 1-bit
ReLU mask instead of keeping the pre-activation.  Numerics are tested against
``torch.nn.functional.batch_norm`` in fp32 (tests/test_gpu_norm.py).
"""
import ctypes

import torch

from spark_deep_learning_b200.ops import comm as _comm
from spark_deep_learning_b200.utils import launch_counter

_DT = {torch.bfloat16: 5, torch.float32: 6}
_sig_done = False
_ws = {}


def _lib():
    global _sig_done
    lib = _comm.load()
    if not _sig_done:
        vp, i32, i64, u64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_float
        lib.sdlc_bn_grid.restype = i32
        lib.sdlc_bn_grid.argtypes = [i64, i32, i32]
        lib.sdlc_bn_fwd.restype = i32
        lib.sdlc_bn_fwd.argtypes = [vp] * 13 + [i64, i32, f32, f32, i32, i32, u64]
        lib.sdlc_bn_bwd.restype = i32
        lib.sdlc_bn_bwd.argtypes = [vp] * 15 + [i64, i32, i32, i32, i32, u64]
        _sig_done = True
    return lib


def _workspace(dev, nfloats):
    key = (dev.index,)
    w = _ws.get(key)
    if w is None or w.numel() < nfloats:
        w = torch.empty(max(nfloats, 1 << 20), dtype=torch.float32, device=dev)
        _ws[key] = w
    return w


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def supported(x):
    """True when the fused sm_100a kernel can take these arguments (device, dtype, layout, alignment)."""
    if not x.is_cuda or x.dim() != 4 or x.dtype not in _DT:
        return False
    if not x.is_contiguous(memory_format=torch.channels_last):
        return False
    c = x.shape[1]
    n = 8 if x.dtype == torch.bfloat16 else 4
    if c % n:
        return False
    cv = c // n
    return (256 % cv == 0) if cv < 256 else (cv % 256 == 0)


class _BNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, momentum, eps, relu, fork=False):
        lib = _lib()
        n, c, h, w = x.shape
        m = n * h * w
        dev = x.device
        y = torch.empty_like(x)                      # preserves channels_last strides
        nvec_elems = 8 if x.dtype == torch.bfloat16 else 4
        mask = torch.empty(m * c // nvec_elems, dtype=torch.uint8, device=dev) if relu else None
        stats = torch.empty(4, c, dtype=torch.float32, device=dev)   # mean, rstd, scale, shift
        grid = lib.sdlc_bn_grid(m, c, _DT[x.dtype])
        ws = _workspace(dev, grid * 2 * c)
        if residual is not None:
            residual = residual.contiguous(memory_format=torch.channels_last)
        rc = lib.sdlc_bn_fwd(_p(x), _p(residual), _p(y), _p(mask), _p(weight), _p(bias), _p(running_mean),
                             _p(running_var), _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]), _p(ws),
                             m, c, float(eps), float(momentum), 1 if relu else 0, _DT[x.dtype],
                             torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError("sdlc_bn_fwd failed (%d)" % rc)
        launch_counter.add(3)
        ctx.save_for_backward(x, mask, weight, bias, stats)
        ctx.set_materialize_grads(False)             # an unused fork handle must not cost a zero tensor
        ctx.relu = relu
        ctx.has_res = residual is not None
        if fork:
            # two handles on the same storage: the consumers' gradients come back separately and are summed inside
            # the backward kernels instead of by an autograd add pass (detach() = same storage, own autograd edge)
            return y, y.detach()
        return y

    @staticmethod
    def backward(ctx, dy, dy2=None):
        lib = _lib()
        x, mask, weight, bias, stats = ctx.saved_tensors
        if dy is None:
            dy, dy2 = dy2, None
        n, c, h, w = x.shape
        m = n * h * w
        dev = x.device
        dy = dy.contiguous(memory_format=torch.channels_last)
        if dy2 is not None:
            dy2 = dy2.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        dz = torch.empty_like(x) if ctx.has_res else None
        red = torch.empty(5, c, dtype=torch.float32, device=dev)     # dweight, dbias, c0, c1, c2
        grid = lib.sdlc_bn_grid(m, c, _DT[x.dtype])
        ws = _workspace(dev, grid * 2 * c)
        # gradient sink (parallel/optimizer.py): the finalize kernel writes dweight / dbias straight into the gradient
        # bucket views, no AccumulateGrad add per parameter
        sinks = None
        if weight is not None and bias is not None:
            sw, sb = getattr(weight, "_sparkdl_grad_sink", None), getattr(bias, "_sparkdl_grad_sink", None)
            if (sw is not None and sb is not None and weight.grad is not None and bias.grad is not None
                    and weight.grad.dtype == torch.float32 and bias.grad.dtype == torch.float32
                    and weight.grad.is_contiguous() and bias.grad.is_contiguous()):
                sinks = (sw, sb)
        dw_t, db_t = (weight.grad, bias.grad) if sinks else (red[0], red[1])
        rc = lib.sdlc_bn_bwd(_p(dy), _p(dy2), _p(x), _p(mask), _p(weight), _p(stats[0]), _p(stats[1]), _p(dx), _p(dz),
                             _p(dw_t), _p(db_t), _p(red[2]), _p(red[3]), _p(red[4]), _p(ws), m, c,
                             1 if ctx.relu else 0, _DT[x.dtype], 1 if sinks else 0,
                             torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError("sdlc_bn_bwd failed (%d)" % rc)
        launch_counter.add(3)
        if sinks:
            sinks[0](weight)
            sinks[1](bias)
            dw = db = None
        else:
            dw = red[0] if weight is not None else None
            db = red[1] if weight is not None else None
        return dx, dw, db, dz, None, None, None, None, None, None


def batch_norm_act(x, bn, residual=None, relu=True, fork=False, count_batches=True):
    """``relu(batch_norm(x) + residual)`` with training-mode statistics, fused.

    ``bn`` is a ``torch.nn.BatchNorm2d`` (its parameters / running stats are used and updated).
    ``fork``: return the result twice (two tensors on one storage) for a consumer pair such as the next residual
    block's first convolution and its identity branch - the two gradients are then added inside the backward kernels.
    ``count_batches=False``: the caller advances ``num_batches_tracked`` itself (one batched op for all layers).
    Falls back to the unfused torch ops for layouts / dtypes the kernels do not cover."""
    if not bn.training or not supported(x) or (residual is not None and residual.dtype != x.dtype):
        if bn.training and not count_batches:
            y = torch.nn.functional.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, True,
                                               bn.momentum if bn.momentum is not None else 0.1, bn.eps)
        else:
            y = bn(x)
        if residual is not None:
            y = y + residual
        y = torch.relu(y) if relu else y
        return (y, y) if fork else y
    momentum = bn.momentum if bn.momentum is not None else 0.1
    if count_batches and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    w = bn.weight.float() if bn.weight is not None and bn.weight.dtype != torch.float32 else bn.weight
    b = bn.bias.float() if bn.bias is not None and bn.bias.dtype != torch.float32 else bn.bias
    return _BNAct.apply(x, w, b, residual, bn.running_mean if bn.track_running_stats else None,
                        bn.running_var if bn.track_running_stats else None, momentum, bn.eps, relu, fork)
