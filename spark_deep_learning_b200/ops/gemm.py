"""bf16 GEMMs on the tcgen05 tensor cores with fused epilogues (``csrc/cuda/gemm_tcgen05.cu``).

One hand-written kernel (CTA-pair tiles, TMA-fed, TMEM accumulators) runs all three GEMMs of a linear layer:

* ``linear_act(x, weight, bias, act)`` == ``act(F.linear(x, weight, bias))``: forward with the bias / GELU epilogue,
* dgrad ``dX = dY . W`` with the GELU derivative fused into the epilogue (no separate elementwise pass),
* wgrad ``dW = dY^T . X`` (split-K over the tokens), optionally accumulated straight into a gradient buffer
  (``wgrad_into``: the K11 path of SURVEY.md §2f - the weight gradient lands in the symmetric fusion bucket that the
  reduce-scatter kernel reads, with no intermediate tensor and no autograd ``add_``),

plus ``conv1x1_stats`` for NHWC 1x1 convolutions whose BatchNorm statistics come out of the GEMM epilogue.
No operand is ever transposed in memory: the kernel takes K-major and MN-major operands.
"""
import ctypes

import torch

from spark_deep_learning_b200.ops import comm as _comm
from spark_deep_learning_b200.utils import launch_counter

ACT = {None: 0, "none": 0, "gelu_tanh": 1, "gelu": 2, "gelu_erf": 2, "deriv": 3}   # "deriv": aux already holds act'(z)
EPI_LINEAR, EPI_DGELU, EPI_ACC_BF16, EPI_ACC_F32, EPI_STATS = range(5)
_sig = False
_NUM_CLUSTERS = 74          # 148 SMs / 2


def _lib():
    global _sig
    lib = _comm.load()
    if not _sig:
        vp, i32, u64, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_longlong
        lib.sdlc_gemm_bf16.restype = i32
        lib.sdlc_gemm_bf16.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, u64]
        lib.sdlc_gemm_ex.restype = i32
        lib.sdlc_gemm_ex.argtypes = [vp, i64, i32, vp, i64, i32, vp, i64, vp, i32, i32, i32, i32, vp, i32, vp, i64,
                                     i32, i32, vp, vp, i32, i32, u64]
        lib.sdlc_gemm_ex3.restype = i32
        lib.sdlc_gemm_ex3.argtypes = [vp, vp, vp, i32, i64, i32, vp, i64, i32, vp, i64, vp, i32, i32, i32, i32, vp, i32, vp,
                                      i64, i32, i32, vp, vp, i32, i32, u64]
        lib.sdlc_gelu_bwd.restype = i32
        lib.sdlc_gelu_bwd.argtypes = [vp, vp, vp, i64, i32, u64]
        lib.sdlc_gemm_stat_rows.restype = i32
        lib.sdlc_gemm_stat_rows.argtypes = [i32]
        lib.sdlc_colsum_bf16.restype = i32
        lib.sdlc_colsum_bf16.argtypes = [vp, i32, i32, i64, vp, i32, i32, vp, vp, u64]
        _sig = True
    return lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


_WS = {}


def _workspace(device, nfloats):
    """Grow-only fp32 scratch for split-K partial tiles (one per device)."""
    ws = _WS.get(device)
    if ws is None or ws.numel() < nfloats:
        ws = torch.empty(max(nfloats, 1 << 22), dtype=torch.float32, device=device)
        _WS[device] = ws
    return ws


def supported(x, weight):
    """True when the fused sm_100a kernel can take these arguments (device, dtype, layout, alignment)."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
            and x.shape[-1] % 8 == 0 and weight.shape[0] % 8 == 0 and weight.shape[1] == x.shape[-1])


def _rowmajor(t):
    """2-D tensor with unit inner stride and a 16-byte aligned pitch; returns (tensor, pitch)."""
    if t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.stride(0) < t.shape[1] or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t, t.stride(0)


def gemm_ex(a, b, m, n, k, a_mn=False, b_mn=False, epi=EPI_ACC_BF16, out=None, out2=None, bias=None, act=None,
            aux=None, accumulate=False, splits=1, stats=None, block_n=0, out2_deriv=False):
    """Raw kernel call: ``out[m,n] = epi(A . B^T)``.  ``a`` is stored [m,k] (K-major) or [k,m] (``a_mn``); ``b`` is
    stored [n,k] or [k,n] (``b_mn``).  ``a`` may also be a tuple of 2-3 equally shaped buffers that together form A -
    split along K for a K-major A, along M for an MN-major A - so that e.g. dQ, dK, dV never have to be concatenated.
    ``out2`` (EPI_LINEAR) receives the pre-activation, or ``act'(pre-activation)`` when ``out2_deriv``.
    Returns ``out``."""
    lib = _lib()
    parts = None
    if isinstance(a, (tuple, list)):
        parts = [_rowmajor(t) for t in a]
        if len({ld for _, ld in parts}) != 1:
            parts = [(t.contiguous(), t.shape[1]) for t, _ in parts]
        a, lda = parts[0]
        part_len = a.shape[1]
        parts = [t for t, _ in parts]
    else:
        a, lda = _rowmajor(a)
    b, ldb = _rowmajor(b)
    dev = a.device
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32 if epi == EPI_ACC_F32 else torch.bfloat16, device=dev)
    ld_out = out.stride(0)
    ws = None
    if splits > 1:
        ws = _workspace(dev, splits * ((m + 255) // 256) * 256 * n)
    act_code = ACT[act] if not isinstance(act, int) else act
    if out2_deriv and out2 is not None and act_code:
        act_code |= 32                           # out2 = act'(z) instead of z
    if bias is not None:
        if bias.dtype == torch.bfloat16 and bias.data_ptr() % 16 == 0:
            act_code |= 16                       # the epilogue reads the bf16 bias directly
        elif bias.dtype != torch.float32:
            bias = bias.float()
    ld_aux = aux.stride(0) if aux is not None else 0
    if parts is not None:
        rc = lib.sdlc_gemm_ex3(_p(parts[0]), _p(parts[1]), _p(parts[2]) if len(parts) > 2 else None, part_len, lda,
                               int(a_mn), _p(b), ldb, int(b_mn), _p(out), ld_out, _p(out2), m, n, k, epi,
                               _p(bias), act_code, _p(aux), ld_aux, int(accumulate),
                               splits, _p(ws), _p(stats), block_n, 0, _stream(a))
    else:
        rc = lib.sdlc_gemm_ex(_p(a), lda, int(a_mn), _p(b), ldb, int(b_mn), _p(out), ld_out, _p(out2), m, n, k, epi,
                              _p(bias), act_code, _p(aux), ld_aux, int(accumulate),
                              splits, _p(ws), _p(stats), block_n, 0, _stream(a))
    if rc != 0:
        raise RuntimeError("sdlc_gemm_ex failed (%d) for M=%d N=%d K=%d epi=%d" % (rc, m, n, k, epi))
    launch_counter.add(1 + (1 if splits > 1 else 0) + (1 if epi == EPI_STATS else 0))
    return out


def gemm_bf16(x2d, weight, bias=None, act=None, save_pre_act=False, block_n=None, save_deriv=False):
    """Forward: x2d [M,K] bf16, weight [N,K] bf16, bias fp32 [N] or None -> ([M,N] bf16, pre-activation or None).
    With ``save_deriv`` the second output is ``act'(pre-activation)`` (what backward needs: pass it to :func:`dgrad` /
    :func:`gelu_bwd` with ``act="deriv"``), computed in the same epilogue where the tanh is already at hand."""
    m, k = x2d.shape
    n = weight.shape[0]
    out = torch.empty((m, n), dtype=torch.bfloat16, device=x2d.device)
    pre = torch.empty_like(out) if save_pre_act else None
    gemm_ex(x2d, weight, m, n, k, epi=EPI_LINEAR, out=out, out2=pre, bias=bias, act=act, block_n=block_n or 0,
            out2_deriv=save_deriv)
    return out, pre


def dgrad(dy2d, weight, pre=None, act=None):
    """``dX[M,K_in] = (dY . W) [* gelu'(pre)]``: dy2d [M,N_out] bf16 (or a tuple of column blocks of it), weight
    [N_out,K_in] bf16 (read as an MN-major B)."""
    m = dy2d[0].shape[0] if isinstance(dy2d, (tuple, list)) else dy2d.shape[0]
    n_out, k_in = weight.shape
    if pre is not None:
        return gemm_ex(dy2d, weight, m, k_in, n_out, b_mn=True, epi=EPI_DGELU, aux=pre, act=act)
    return gemm_ex(dy2d, weight, m, k_in, n_out, b_mn=True, epi=EPI_ACC_BF16)


def _wgrad_splits(n_out, k_in, tokens):
    tiles = ((n_out + 255) // 256) * ((k_in + 255) // 256)
    kb = (tokens + 63) // 64
    s = max(1, _NUM_CLUSTERS // max(tiles, 1))
    return max(1, min(s, kb // 4, 16))


def wgrad(dy2d, x2d, out=None, accumulate=False):
    """``dW[N_out,K_in] (+)= dY^T . X`` without transposing anything: both operands are read MN-major.  ``out`` may be
    bf16 or fp32 (e.g. a view of a gradient bucket)."""
    if isinstance(dy2d, (tuple, list)):
        tokens, n_out = dy2d[0].shape[0], sum(t.shape[1] for t in dy2d)
    else:
        tokens, n_out = dy2d.shape
    k_in = x2d.shape[1]
    if out is None:
        out = torch.empty((n_out, k_in), dtype=torch.bfloat16, device=x2d.device)
    epi = EPI_ACC_F32 if out.dtype == torch.float32 else EPI_ACC_BF16
    return gemm_ex(dy2d, x2d, n_out, k_in, tokens, a_mn=True, b_mn=True, epi=epi, out=out, accumulate=accumulate,
                   splits=_wgrad_splits(n_out, k_in, tokens))


def conv1x1_stats(x2d, weight):
    """NHWC 1x1 convolution as a GEMM (``x2d`` [N*H*W, Cin], ``weight`` [Cout, Cin]) whose epilogue also produces
    the per-channel sum and sum of squares of the bf16 output.  Returns (y [M,Cout] bf16, totals fp32 [2,Cout])."""
    m, k = x2d.shape
    n = weight.shape[0]
    rows = _lib().sdlc_gemm_stat_rows(m)
    stats = torch.empty((2 * rows + 2) * n, dtype=torch.float32, device=x2d.device)
    y = gemm_ex(x2d, weight, m, n, k, epi=EPI_STATS, stats=stats)
    return y, stats[2 * rows * n:].view(2, n)


_CS = {}


def colsum(x2d, out=None, accumulate=False):
    """``out[N] (+)= x2d.sum(0)`` for a bf16 matrix, deterministic, one launch (the bias gradient of a linear layer).
    ``out`` may be fp32 or bf16 (e.g. the view of a bias inside a gradient bucket)."""
    m, n = x2d.shape
    x2d, ld = _rowmajor(x2d)
    dev = x2d.device
    ws = _CS.get(dev)
    npad = (n + 63) // 64 * 64
    if ws is None or ws[0].numel() < 64 * npad:
        ws = (torch.empty(64 * max(npad, 32768), dtype=torch.float32, device=dev),
              torch.zeros(max(npad, 32768) // 64, dtype=torch.int32, device=dev))
        _CS[dev] = ws
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=dev)
    rc = _lib().sdlc_colsum_bf16(_p(x2d), m, n, ld, _p(out), 1 if out.dtype == torch.float32 else 0, int(accumulate),
                                 _p(ws[0]), _p(ws[1]), _stream(x2d))
    if rc != 0:
        raise RuntimeError("sdlc_colsum_bf16 failed (%d)" % rc)
    launch_counter.add(1)
    return out


def _bias_grad(dy2d, bias, use_sink=True):
    """Gradient of a bias: accumulated straight into ``bias.grad`` (bucket view) when the optimizer offers a sink."""
    sink = _grad_sink(bias) if use_sink else None
    if sink is not None and bias.grad.dtype in (torch.float32, torch.bfloat16) and dy2d.shape[1] % 8 == 0:
        colsum(dy2d, out=bias.grad, accumulate=True)
        sink(bias)
        return None
    if dy2d.shape[1] % 8 == 0 and dy2d.dtype == torch.bfloat16:
        return colsum(dy2d)
    return dy2d.sum(0, dtype=torch.float32)


def gelu_bwd(dy2d, pre, act):
    """``dy2d * gelu'(pre)`` as a stand-alone pass (when the product is needed as a GEMM operand); ``act="deriv"``:
    ``pre`` already holds the derivative."""
    g = torch.empty_like(pre)
    rc = _lib().sdlc_gelu_bwd(_p(dy2d.contiguous()), _p(pre), _p(g), pre.numel(), ACT[act], _stream(pre))
    if rc != 0:
        raise RuntimeError("sdlc_gelu_bwd failed")
    launch_counter.add(1)
    return g


def _grad_sink(weight):
    """The buffer the weight gradient should be accumulated into directly, or None.

    ``DistributedOptimizer`` keeps ``weight.grad`` as a view of its symmetric gradient bucket and marks the parameter
    with ``_sparkdl_grad_sink`` (a callable ``(param) -> None`` that plays the role of the post-accumulate hook).
    """
    sink = getattr(weight, "_sparkdl_grad_sink", None)
    if sink is None or weight.grad is None or not weight.grad.is_contiguous():
        return None
    return sink


class _LinearAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act, use_sink=True):
        ctx.use_sink = use_sink
        x2d = x.reshape(-1, x.shape[-1])
        need_pre = act not in (None, "none") and (x.requires_grad or weight.requires_grad)
        out, pre = gemm_bf16(x2d, weight, bias, act, save_pre_act=need_pre, save_deriv=True)   # pre = act'(z)
        ctx.save_for_backward(x2d, weight, pre, bias)
        ctx.act = "deriv"
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return out.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2d, weight, pre, bias = ctx.saved_tensors
        dy2d = dy.reshape(-1, dy.shape[-1])
        dy2d, _ = _rowmajor(dy2d)
        need_dx, need_dw, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        if pre is not None and (need_dw or need_db):
            # dY * gelu'(pre) is needed as an operand of wgrad as well: materialise it once
            dy2d = gelu_bwd(dy2d, pre, ctx.act)
            pre = None
        dx = dgrad(dy2d, weight, pre, ctx.act).view(ctx.x_shape) if need_dx else None
        dw = None
        if need_dw:
            sink = _grad_sink(weight) if ctx.use_sink else None
            if sink is not None:
                # K11: the wgrad epilogue accumulates straight into the gradient bucket; autograd sees no tensor
                wgrad(dy2d, x2d, out=weight.grad, accumulate=True)
                sink(weight)
            else:
                dw = wgrad(dy2d, x2d)
        db = _bias_grad(dy2d, bias, ctx.use_sink) if need_db else None
        return dx, dw, db, None, None


class _QKVProj(torch.autograd.Function):
    """Packed QKV projection that hands attention three views of one [M, 3H] buffer and takes dQ, dK, dV back as three
    separate buffers: the dgrad (split along K) and wgrad (split along M) GEMMs read them in place, so the
    ``stack`` + ``permute`` copies that autograd inserts for a packed projection disappear."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2d = x.reshape(-1, x.shape[-1])
        qkv, _ = gemm_bf16(x2d, weight, bias, None)
        h = weight.shape[0] // 3
        ctx.save_for_backward(x2d, weight, bias)
        ctx.x_shape = x.shape
        lead = x.shape[:-1]
        return tuple(qkv[:, i * h:(i + 1) * h].view(*lead, h) for i in range(3))

    @staticmethod
    def backward(ctx, dq, dk, dv):
        x2d, weight, bias = ctx.saved_tensors
        h = weight.shape[0] // 3
        parts = []
        for g in (dq, dk, dv):
            g2 = g.reshape(-1, h)
            if g2.stride(1) != 1 or g2.stride(0) != h or g2.data_ptr() % 16:
                g2 = g2.contiguous()
            parts.append(g2)
        dx = dgrad(parts, weight).view(ctx.x_shape) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1]:
            sink = _grad_sink(weight)
            if sink is not None:
                wgrad(parts, x2d, out=weight.grad, accumulate=True)
                sink(weight)
            else:
                dw = wgrad(parts, x2d)
        if bias is not None and ctx.needs_input_grad[2]:
            sink = _grad_sink(bias)
            if sink is not None and bias.grad.dtype in (torch.float32, torch.bfloat16):
                for i, g2 in enumerate(parts):
                    colsum(g2, out=bias.grad[i * h:(i + 1) * h], accumulate=True)
                sink(bias)
            else:
                db = torch.cat([colsum(g2) for g2 in parts])
        return dx, dw, db


def qkv_proj(x, weight, bias):
    """``(q, k, v) = split(x @ weight.T + bias, 3)`` for a packed [3H, H] projection (bf16, tcgen05 GEMMs)."""
    if not (supported(x, weight) and weight.shape[0] % (3 * 256) == 0):
        y = torch.nn.functional.linear(x, weight, None if bias is None else bias.to(x.dtype))
        return y.chunk(3, dim=-1)
    return _QKVProj.apply(x, weight, bias)


class _MLP(torch.autograd.Function):
    """``act(x W1^T + b1) W2^T + b2`` as one autograd node, so that the GELU derivative is applied inside the epilogue
    of the second layer's dgrad GEMM (``dZ = (dY . W2) * gelu'(Z)``) instead of a separate elementwise pass."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, act):
        x2d = x.reshape(-1, x.shape[-1])
        h, z = gemm_bf16(x2d, w1, b1, act, save_pre_act=True, save_deriv=True)     # z = act'(x W1^T + b1)
        y, _ = gemm_bf16(h, w2, b2, None)
        ctx.save_for_backward(x2d, w1, w2, z, h, b1, b2)
        ctx.act = "deriv"
        ctx.x_shape = x.shape
        ctx.has_b = (b1 is not None, b2 is not None)
        return y.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2d, w1, w2, z, h, b1, b2 = ctx.saved_tensors
        dy2d, _ = _rowmajor(dy.reshape(-1, dy.shape[-1]))
        dz = dgrad(dy2d, w2, z, ctx.act)                     # fused: (dY . W2) * gelu'(Z), gelu'(Z) saved by forward
        grads = {}
        for name, w, g, inp in (("w2", w2, dy2d, h), ("w1", w1, dz, x2d)):
            sink = _grad_sink(w)
            if sink is not None:
                wgrad(g, inp, out=w.grad, accumulate=True)
                sink(w)
                grads[name] = None
            else:
                grads[name] = wgrad(g, inp)
        db2 = _bias_grad(dy2d, b2) if ctx.has_b[1] else None
        db1 = _bias_grad(dz, b1) if ctx.has_b[0] else None
        dx = dgrad(dz, w1).view(ctx.x_shape) if ctx.needs_input_grad[0] else None
        return dx, grads["w1"], db1, grads["w2"], db2, None


def mlp(x, w1, b1, w2, b2, act="gelu"):
    """Transformer feed-forward block ``act(x W1^T + b1) W2^T + b2`` (bf16), six tcgen05 GEMMs per fwd+bwd."""
    if not (supported(x, w1) and w2.dtype == torch.bfloat16 and w2.shape[0] % 8 == 0 and w1.shape[0] % 8 == 0):
        return linear_act(linear_act(x, w1, b1, act), w2, b2, None)
    return _MLP.apply(x, w1, b1, w2, b2, act)


def linear_act(x, weight, bias=None, act=None, direct_wgrad=True):
    """``act(x @ weight.T + bias)``; bf16 in / out, fp32 accumulate, all three GEMMs on the tcgen05 kernel.

    ``direct_wgrad=False`` keeps the weight gradient on the autograd path (needed when ``weight`` also receives
    gradient from another use in the same backward, e.g. a tied embedding / decoder matrix)."""
    if not supported(x, weight):
        y = torch.nn.functional.linear(x, weight, None if bias is None else bias.to(x.dtype))
        if act in ("gelu", "gelu_erf"):
            return torch.nn.functional.gelu(y)
        if act == "gelu_tanh":
            return torch.nn.functional.gelu(y, approximate="tanh")
        return y
    return _LinearAct.apply(x, weight, bias, act, direct_wgrad)
