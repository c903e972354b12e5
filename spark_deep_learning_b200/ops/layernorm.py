"""Fused ``LayerNorm(x + dropout(a))`` (``csrc/cuda/ln_fused.cu``): one streaming pass forward, one backward, bf16
activations, fp32 statistics, 1-bit dropout mask, graph-safe Philox counters kept in device memory."""
import ctypes
import itertools

import torch

from spark_deep_learning_b200.ops import comm as _comm
from spark_deep_learning_b200.utils import launch_counter

_sig = False
_rng = {}
_ws = {}
_call_ids = itertools.count(1)


def _lib():
    global _sig
    lib = _comm.load()
    if not _sig:
        vp, i32, u32, u64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_ulonglong, ctypes.c_float
        lib.sdlc_ln_supported.restype = i32
        lib.sdlc_ln_supported.argtypes = [i32]
        lib.sdlc_ln_grid.restype = i32
        lib.sdlc_ln_grid.argtypes = [i32]
        lib.sdlc_ln_fwd.restype = i32
        lib.sdlc_ln_fwd.argtypes = [vp] * 10 + [u32, i32, i32, f32, f32, u64]
        lib.sdlc_ln_bwd.restype = i32
        lib.sdlc_ln_bwd.argtypes = [vp] * 11 + [i32, i32, f32, u64]
        _sig = True
    return lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _rng_state(dev):
    st = _rng.get(dev.index)
    if st is None:
        st = torch.tensor([torch.initial_seed() & 0x7fffffffffffffff, 0], dtype=torch.int64, device=dev)
        _rng[dev.index] = st
    return st


def _workspace(dev, n):
    w = _ws.get(dev.index)
    if w is None or w.numel() < n:
        w = torch.empty(max(n, 1 << 20), dtype=torch.float32, device=dev)
        _ws[dev.index] = w
    return w


def supported(x, weight):
    """True when the fused sm_100a kernel can take these arguments (device, dtype, layout, alignment)."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x.is_contiguous()
            and _lib().sdlc_ln_supported(x.shape[-1]) == 1)


class _DropAddLN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a, weight, bias, p, eps, call_id):
        lib = _lib()
        h = x.shape[-1]
        t = x.numel() // h
        dev = x.device
        y = torch.empty_like(x)
        z = torch.empty_like(x) if a is not None else None
        use_mask = a is not None and p > 0.0
        mask = torch.empty(t * h // 8, dtype=torch.uint8, device=dev) if use_mask else None
        stats = torch.empty(2, t, dtype=torch.float32, device=dev)
        rc = lib.sdlc_ln_fwd(_p(x), _p(a), _p(weight), _p(bias), _p(y), _p(z), _p(mask), _p(stats[0]), _p(stats[1]),
                             _p(_rng_state(dev)), call_id, t, h, float(p), float(eps),
                             torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError("sdlc_ln_fwd failed (%d)" % rc)
        launch_counter.add(2 if use_mask else 1)
        ctx.save_for_backward(z if z is not None else x, mask, weight, stats)
        ctx.p = float(p) if use_mask else 0.0
        ctx.has_a = a is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib()
        z, mask, weight, stats = ctx.saved_tensors
        h = z.shape[-1]
        t = z.numel() // h
        dev = z.device
        dy = dy.contiguous()
        dx = torch.empty_like(z)
        da = torch.empty_like(z) if ctx.has_a else None
        grid = lib.sdlc_ln_grid(t)
        ws = _workspace(dev, grid * 2 * h)
        dgb = torch.empty(2, h, dtype=torch.float32, device=dev)
        rc = lib.sdlc_ln_bwd(_p(dy), _p(z), _p(weight), _p(mask), _p(stats[0]), _p(stats[1]), _p(dx), _p(da), _p(ws),
                             _p(dgb[0]), _p(dgb[1]), t, h, ctx.p, torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError("sdlc_ln_bwd failed (%d)" % rc)
        launch_counter.add(2)
        return dx, da, dgb[0].to(weight.dtype), dgb[1].to(weight.dtype), None, None, None


class FusedDropoutAddLayerNorm(torch.nn.Module):
    """``LayerNorm(x + dropout(a))`` (``a=None``: plain LayerNorm).  Falls back to torch ops when unsupported."""

    def __init__(self, hidden, p=0.0, eps=1e-12, fused=True):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.ones(hidden))
        self.bias = torch.nn.Parameter(torch.zeros(hidden))
        self.p, self.eps, self.fused = p, eps, fused
        self._call_id = next(_call_ids)

    def forward(self, x, a=None):
        p = self.p if self.training else 0.0
        if self.fused and supported(x, self.weight) and (a is None or (a.dtype == x.dtype and a.shape == x.shape)):
            return _DropAddLN.apply(x, a.contiguous() if a is not None else None, self.weight, self.bias, p, self.eps,
                                    self._call_id)
        if a is not None:
            x = x + torch.nn.functional.dropout(a, p, self.training)
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), self.weight.to(x.dtype), self.bias.to(x.dtype), self.eps)
