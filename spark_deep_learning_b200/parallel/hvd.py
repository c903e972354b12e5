"""Horovod-style collective frontend used inside a pickled ``main``.

The reference *is* "run Horovod" (``/root/reference/sparkdl/horovod/runner_base.py:24-25``) but
ships no frontend: ``hvd.init()`` in the user's main is expected to find its peers
(SURVEY.md §1b "Rendezvous + Horovod runtime").  This module is that frontend, written from
scratch for one B200 box:

* CPU tensors  -> C++ engine + POSIX-shm collectives (``csrc/core``);
* CUDA tensors -> the same engine negotiates names / fusion, then launches hand-written
  sm_100a peer-memory kernels (``csrc/cuda``) on a communication stream - no NCCL call.

Usage (same names as ``horovod.torch``)::

    import sparkdl.hvd as hvd
    hvd.init()
    opt = hvd.DistributedOptimizer(opt, named_parameters=model.named_parameters())
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
"""
import io
import os
import pickle
import threading
import weakref
import uuid

import torch

from spark_deep_learning_b200.parallel import context as _context
from spark_deep_learning_b200.parallel import core as _core
from spark_deep_learning_b200.utils.config import get_config

# reduce-op constants (values match csrc/core/common.h ReduceOp where applicable)
Average = _core.OP_AVERAGE
Sum = _core.OP_SUM
Min = _core.OP_MIN
Max = _core.OP_MAX
Product = _core.OP_PRODUCT
Adasum = 100

_DTYPES = {
    torch.uint8: _core.U8, torch.int8: _core.I8, torch.int32: _core.I32, torch.int64: _core.I64,
    torch.float16: _core.F16, torch.bfloat16: _core.BF16, torch.float32: _core.F32,
    torch.float64: _core.F64, torch.bool: _core.BOOL,
}
_FLOATS = (torch.float16, torch.bfloat16, torch.float32, torch.float64)


class HorovodInternalError(RuntimeError):
    """A collective failed (mismatched submissions, a dead peer, a stall)."""


class _State(object):
    def __init__(self):
        self.initialized = False
        self.ctx = None
        self.cfg = None
        self.engine = None
        self.comm = None          # ops.comm.SymmComm when CUDA is in use
        self.counters = {}
        self.lock = threading.Lock()
        self.inflight = {}        # key -> (weakref to the handle, engine handles, tensors kept alive, is_cuda)
        self.inflight_seq = 0
        self.inflight_purge_at = 64


_S = _State()


# ---------------------------------------------------------------------------------------------
# lifecycle
# ---------------------------------------------------------------------------------------------
def init(comm=None, process_sets=None):  # pylint: disable=unused-argument
    """Join the gang started by :class:`sparkdl.HorovodRunner` (or torchrun) and start the engine.

    On a GPU box this also pins the process to ``cuda:local_rank`` and exchanges CUDA VMM/IPC
    handles, signal pads and the NVSwitch multicast object with the peers
    (BASELINE.json north star: "exchanges cudaIpc/fabric handles and NVLink topology at init").
    """
    with _S.lock:
        if _S.initialized:
            return
        cfg = get_config(refresh=True)
        ctx = _context.get_context()
        # re-init support (Horovod allows shutdown() -> init()): every incarnation talks to the store through its
        # own key namespace, so stale bootstrap keys of the previous one are never observed.
        _S.generation = getattr(_S, "generation", 0) + 1
        if getattr(ctx, "base_store", None) is None:
            ctx.base_store = ctx.store
        ctx.store = _context.NamespacedStore(ctx.base_store, "g%d/" % _S.generation)
        use_cuda = torch.cuda.is_available() and os.environ.get("SPARKDL_FORCE_CPU") != "1"
        if use_cuda:
            torch.cuda.set_device(ctx.local_rank % torch.cuda.device_count())
            from spark_deep_learning_b200.utils import affinity
            _S.affinity = affinity.bind_to_gpu(torch.cuda.current_device())   # external gangs (torchrun) get it here
        if ctx.rank == 0:
            shm_name = "/sparkdl-%s-%s" % (ctx.job_id[:40], uuid.uuid4().hex[:8])
            ctx.store.set("shm_name", shm_name)
        else:
            shm_name = ctx.store.get("shm_name", cfg.start_timeout_s)
        _S.engine = _core.CoreEngine(shm_name, ctx.rank, ctx.size, cfg)
        _S.engine.set_cache_enabled(bool(getattr(cfg, "response_cache", True)))
        _S.ctx, _S.cfg = ctx, cfg
        _S.counters = {}
        if use_cuda:
            from spark_deep_learning_b200.ops import comm as _comm
            _S.comm = _comm.SymmComm(ctx, cfg)
            fn, cctx = _S.comm.engine_executor()
            _S.engine.set_gpu_executor(fn, cctx)
        _S.initialized = True
        if not getattr(_S, "atexit_registered", False):
            import atexit
            atexit.register(_shutdown_at_exit)
            _S.atexit_registered = True


def _shutdown_at_exit():
    """A single-process script that forgets ``hvd.shutdown()`` would leave its 27 MB shared-memory segment behind.
    Gangs are left alone here: their teardown is collective and is driven by the launcher / the user's code."""
    try:
        if _S.initialized and _S.ctx is not None and _S.ctx.size == 1:
            shutdown()
    except Exception:  # pylint: disable=broad-except
        pass


def shutdown():
    """Drain outstanding operations, stop the engine thread and release the symmetric buffers (idempotent)."""
    with _S.lock:
        if not _S.initialized:
            return
        _S.initialized = False
        try:
            _purge_inflight(wait=True)       # operations whose handles were dropped still own their tensors
        except Exception:  # pylint: disable=broad-except
            pass
        try:
            if _S.comm is not None:
                _S.comm.close()
        finally:
            _S.comm = None
            try:
                _S.engine.shutdown()
            finally:
                _S.engine = None
                _S.inflight.clear()          # only now: the engine thread is gone, nobody writes the tensors any more
                _S.ctx = None   # the job context (store connection) stays alive for a later init()


def rank_hint():
    """Rank of this process as the launcher assigned it; usable before ``init()`` / after ``shutdown()``."""
    return _context.get_context().rank


def is_initialized():
    """True between ``init()`` and ``shutdown()``."""
    return _S.initialized


def _require():
    if not _S.initialized:
        raise ValueError("sparkdl.hvd has not been initialized; use hvd.init().")


def size():
    """Number of workers in the gang (``np`` of ``HorovodRunner``)."""
    _require()
    return _S.ctx.size


def rank():
    """This worker's rank, ``0 .. size()-1``."""
    _require()
    return _S.ctx.rank


def local_rank():
    """Rank among the workers of this host; equals the GPU index the worker is pinned to."""
    _require()
    return _S.ctx.local_rank


def local_size():
    """Number of workers on this host."""
    _require()
    return _S.ctx.local_size


def cross_rank():
    """Index of this host among the hosts of the job (always 0: single-box runtime)."""
    _require()
    return 0


def cross_size():
    """Number of hosts in the job (always 1: single-box runtime)."""
    _require()
    return 1


def is_homogeneous():
    """True when every host runs the same number of workers (trivially true on one box)."""
    return True


def mpi_threads_supported():
    """Horovod compatibility probe: there is no MPI in this runtime."""
    return False


def mpi_enabled():
    """Horovod compatibility probe: there is no MPI in this runtime."""
    return False


def mpi_built():
    """Horovod compatibility probe: there is no MPI in this runtime."""
    return False


def gloo_enabled():
    """Horovod compatibility probe: the controller is the shm engine, not Gloo."""
    return False


def gloo_built():
    """Horovod compatibility probe: the controller is the shm engine, not Gloo."""
    return False


def nccl_built():
    """Horovod compatibility probe.  GPU collectives are this package's own sm_100a kernels, not NCCL."""
    return 0  # by design: the data plane is our own sm_100a kernels, not NCCL


def ddl_built():
    """Horovod compatibility probe (IBM DDL): always False."""
    return False


def ccl_built():
    """Horovod compatibility probe (oneCCL): always False."""
    return False


def cuda_built():
    """True when a CUDA device is usable by this process."""
    return torch.cuda.is_available()


def rocm_built():
    """Horovod compatibility probe: always False."""
    return False


def comm_backend():
    """The CUDA symmetric-memory communicator (``None`` on a CPU-only gang)."""
    _require()
    return _S.comm


def engine_stats():
    """Counters of the C++ engine as a dict: ``responses`` executed, ``tensors`` fused into them, response-cache
    ``cache_hits`` / ``cache_misses`` / ``cache_size`` (collectives announced with one bit instead of a request record)
    and ``negotiation_us`` (mean time from ``enqueue`` to the agreed response on this rank)."""
    _require()
    return _S.engine.stats()


class ProcessSet(object):
    """Horovod's process-set handle.  One box, one gang: only the global set exists here."""

    def __init__(self, ranks=None):
        self.ranks = None if ranks is None else list(ranks)
        self.process_set_id = 0 if ranks is None else None

    def size(self):
        return size() if self.ranks is None else len(self.ranks)

    def rank(self):
        return rank() if self.ranks is None else self.ranks.index(rank())

    def included(self):
        return self.ranks is None or (is_initialized() and rank() in self.ranks)


global_process_set = ProcessSet()


def add_process_set(process_set):
    """Sub-communicators are not part of the reference's contract (its only axis is ``np``); refuse loudly."""
    ranks = process_set.ranks if isinstance(process_set, ProcessSet) else list(process_set)
    if is_initialized() and sorted(ranks) == list(range(size())):
        return global_process_set
    raise NotImplementedError("only the global process set is supported by the single-box sm_100a runtime")


def remove_process_set(process_set):  # pylint: disable=unused-argument
    """Process sets other than the global one are not supported; mirrors Horovod's signature."""
    return False


def start_timeline(file_path, mark_cycles=False):  # pylint: disable=unused-argument
    """Start writing the Chrome-trace timeline of this rank to ``file_path`` (rank r > 0 appends ``.r``).

    Same effect as ``HOROVOD_TIMELINE`` / ``SPARKDL_TIMELINE`` at ``hvd.init()``, but switchable at run time.
    Raises ``ValueError`` when a timeline is already being recorded or the file cannot be created."""
    _require()
    if not _S.engine.timeline_start(str(file_path)):
        raise ValueError("timeline is already active or %r cannot be opened for writing" % (file_path,))


def stop_timeline():
    """Stop the chrome-trace timeline started by ``start_timeline`` / ``SPARKDL_TIMELINE`` and close the file."""
    _require()
    _S.engine.timeline_stop()


# ---------------------------------------------------------------------------------------------
# compression (wire dtype != storage dtype)
# ---------------------------------------------------------------------------------------------
class _NoneCompressor(object):
    @staticmethod
    def compress(tensor):
        return tensor, None

    @staticmethod
    def decompress(tensor, ctx):  # pylint: disable=unused-argument
        return tensor


class _CastCompressor(object):
    wire = torch.float16

    @classmethod
    def compress(cls, tensor):
        if tensor.dtype in (torch.float32, torch.float64):
            return tensor.to(cls.wire), tensor.dtype
        return tensor, None

    @staticmethod
    def decompress(tensor, ctx):
        return tensor if ctx is None else tensor.to(ctx)


class _FP16Compressor(_CastCompressor):
    wire = torch.float16


class _BF16Compressor(_CastCompressor):
    wire = torch.bfloat16


class Compression(object):
    """Gradient compression for allreduce (``hvd.Compression.fp16`` as in Horovod)."""
    none = _NoneCompressor
    fp16 = _FP16Compressor
    bf16 = _BF16Compressor


# ---------------------------------------------------------------------------------------------
# handles
# ---------------------------------------------------------------------------------------------
class _Handle(object):
    __slots__ = ("hs", "output", "keep", "post", "cuda", "done", "result", "key", "__weakref__")

    def __init__(self, hs, output, keep, post=None, cuda=False):
        self.hs = hs
        self.output = output
        self.keep = keep
        self.post = post
        self.cuda = cuda
        self.done = False
        self.result = None
        self.key = None
        if hs:
            _track(self)


# ---------------------------------------------------------------------------------------------
# In-flight registry.  The engine thread reads the input and writes the output of an asynchronous operation until it
# completes, whatever the caller does with the returned handle.  A handle that is dropped without synchronize() (or
# overwritten by the next iteration's list) must therefore not take its tensors with it: the registry keeps them alive
# until the engine reports completion, and frees the engine-side state of handles nobody will synchronize any more.
# (Found with a CPU soak test: 64 dropped allreduce handles -> the engine wrote into freed tensors -> heap corruption.)
# ---------------------------------------------------------------------------------------------
def _track(handle):
    reg = _S.inflight
    _S.inflight_seq += 1
    handle.key = _S.inflight_seq             # not id(): ids are recycled while the old operation may still be running
    reg[handle.key] = (weakref.ref(handle), tuple(handle.hs), handle.keep, handle.cuda)
    if len(reg) >= _S.inflight_purge_at:
        _purge_inflight()
        _S.inflight_purge_at = max(64, 2 * len(reg))


def _purge_inflight(wait=False):
    """Forget completed operations; release those whose handle object is gone.  ``wait``: give each up to 5 s to
    complete first (shutdown path); entries that still have not completed stay registered."""
    eng = _S.engine
    if eng is None:
        return
    patience = wait
    for key, (ref, hs, _keep, cuda) in list(_S.inflight.items()):
        if patience:
            for h in hs:
                try:
                    eng.wait(h, timeout=5.0)
                except TimeoutError:
                    patience = False         # a peer is gone: do not wait for the rest either
                except Exception:  # pylint: disable=broad-except
                    pass
        if not all(eng.poll(h) != 0 for h in hs):
            continue
        if ref() is None:                    # dropped by the caller: nobody else will release it
            for h in hs:
                if cuda and _S.comm is not None:
                    try:
                        ev = eng.done_event(h)
                        if ev:
                            _S.comm.current_stream_wait(ev, release=True)   # order the buffers' reuse after the kernel
                    except Exception:  # pylint: disable=broad-except
                        pass
                eng.release(h)
        del _S.inflight[key]


def _auto_name(kind, name):
    if name is not None:
        full = "%s.%s" % (kind, name)
        if len(full.encode()) > 120:
            # the engine's request slots hold 127 name bytes: keep long names unique with a digest suffix
            import hashlib
            full = full.encode()[:96].decode("utf-8", "ignore") + "#" + hashlib.sha1(full.encode()).hexdigest()[:20]
        return full
    n = _S.counters.get(kind, 0)
    _S.counters[kind] = n + 1
    return "%s.noname.%d" % (kind, n)


def _dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise ValueError("sparkdl.hvd does not support tensors of dtype %s" % t.dtype)


def _device_of(t):
    return t.device.index if t.is_cuda else -1


def _ready_event(t):
    """CUDA tensors: event recorded on the producer stream so the comm stream can order itself
    after the kernels that wrote ``t``."""
    if not t.is_cuda:
        return 0
    if _S.comm is None:
        raise ValueError("a CUDA tensor was submitted but this gang runs the CPU-only runtime "
                         "(no GPU visible at hvd.init(), or SPARKDL_FORCE_CPU=1)")
    return _S.comm.record_ready_event()


_FAULT = {"parsed": False, "kind": None, "rank": -1, "op": -1, "count": 0}


def _maybe_fault():
    """``SPARKDL_FAULT=kill:rank=R:op=K`` / ``raise:rank=R:op=K``: fail rank R at its K-th collective
    (SURVEY.md §4b T5 - failure detection is tested by injecting the failure mid-training)."""
    f = _FAULT
    if not f["parsed"]:
        f["parsed"] = True
        spec = os.environ.get("SPARKDL_FAULT", "")
        parts = spec.split(":")
        opts = dict(p.split("=", 1) for p in parts[1:] if "=" in p)
        if parts[0] in ("kill", "raise") and "op" in opts:
            f.update(kind=parts[0], rank=int(opts.get("rank", -1)), op=int(opts["op"]))
    if f["kind"] is None or f["rank"] != _S.ctx.rank:
        return
    f["count"] += 1
    if f["count"] == f["op"]:
        if f["kind"] == "kill":
            os._exit(137)
        raise RuntimeError("injected fault on rank %d at collective #%d" % (f["rank"], f["op"]))


def _enqueue(req, name, t_in, t_out, op=_core.OP_SUM, root=0, pre=1.0, post=1.0, splits=None,
             shape=None):
    _maybe_fault()
    return _S.engine.enqueue(req, name, _dtype_code(t_in), _device_of(t_in), op, root,
                             list(t_in.shape if shape is None else shape), pre, post,
                             t_in.data_ptr(), t_out.data_ptr() if t_out is not None else 0,
                             splits, _ready_event(t_in))


def poll(handle):
    """True when :func:`synchronize` would not block on the host."""
    _require()
    if handle.done:
        return True
    if not handle.hs and isinstance(handle.keep, list):        # grouped handle: ready when every member is
        return all(poll(sub) for sub in handle.keep)
    return all(_S.engine.poll(h) != 0 for h in handle.hs)


def synchronize(handle):
    """Wait for an asynchronous collective and return its output tensor.

    For CUDA tensors the host only waits until the kernel has been *launched*; the calling
    stream is then made to wait for the completion event, so no host-device sync happens.
    """
    _require()
    if handle.done:
        return handle.result
    _S.inflight.pop(handle.key, None)
    try:
        for h in handle.hs:
            try:
                _S.engine.wait(h)
            except _core.EngineError as e:
                raise HorovodInternalError(str(e))
            if handle.cuda:
                ev = _S.engine.done_event(h)
                if ev:
                    _S.comm.current_stream_wait(ev, release=True)
        if handle.cuda:
            # a peer that stopped after negotiation makes the kernel's flag barrier time out; surface that
            # instead of returning partial data
            try:
                _S.comm.check_error()
            except Exception as e:  # pylint: disable=broad-except
                raise HorovodInternalError(str(e))
        out = handle.output
        if handle.post is not None:
            out = handle.post(handle)
    finally:
        for h in handle.hs:
            _S.engine.release(h)
    handle.done = True
    handle.result = out
    handle.keep = None
    return out


# ---------------------------------------------------------------------------------------------
# allreduce
# ---------------------------------------------------------------------------------------------
def _resolve_op(average, op):
    if average is not None and op is not None:
        raise ValueError("The op parameter supersedes average. Please provide only one of them.")
    if op is None:
        op = Average if (average is None or average) else Sum
    if op == Adasum:
        raise NotImplementedError("Adasum is not supported by the sm_100a runtime")
    return op


def _allreduce_async_impl(tensor, output, name, op, prescale, postscale):
    _require()
    if op == Average:
        if tensor.dtype in _FLOATS:
            op, postscale = Sum, postscale / float(size())
            int_avg = False
        else:
            op, int_avg = Sum, True
    else:
        int_avg = False
    src = tensor if tensor.is_contiguous() else tensor.contiguous()
    dst = output if (output.is_contiguous() and output.dtype == src.dtype) else torch.empty_like(src)
    h = _enqueue(_core.ALLREDUCE, name, src, dst, op=op, pre=prescale, post=postscale)

    def post(hd, _dst=dst, _out=output, _int_avg=int_avg):
        if _int_avg:
            _dst.copy_(torch.div(_dst, size(), rounding_mode="floor"))
        if _dst is not _out:
            _out.copy_(_dst)
        return _out
    return _Handle([h], output, (src, dst, tensor), post, tensor.is_cuda)


def allreduce_async(tensor, average=None, name=None, op=None, prescale_factor=1.0,
                    postscale_factor=1.0, process_set=None):  # pylint: disable=unused-argument
    """Start an out-of-place allreduce; returns a handle for ``synchronize`` / ``poll``."""
    op = _resolve_op(average, op)
    output = torch.empty_like(tensor, memory_format=torch.contiguous_format)
    return _allreduce_async_impl(tensor, output, _auto_name("allreduce", name), op,
                                 prescale_factor, postscale_factor)


def allreduce_async_(tensor, average=None, name=None, op=None, prescale_factor=1.0,
                     postscale_factor=1.0, process_set=None):  # pylint: disable=unused-argument
    """Start an in-place allreduce of ``tensor``; returns a handle for ``synchronize`` / ``poll``."""
    op = _resolve_op(average, op)
    return _allreduce_async_impl(tensor, tensor, _auto_name("allreduce", name), op,
                                 prescale_factor, postscale_factor)


class _AllreduceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, average, name, op, pre, post):
        ctx.args = (average, op, pre, post)
        return synchronize(allreduce_async(tensor, average, name, op, pre, post))

    @staticmethod
    def backward(ctx, grad):
        average, op, pre, post = ctx.args
        return allreduce(grad.contiguous(), average=average, op=op, prescale_factor=pre,
                         postscale_factor=post), None, None, None, None, None


def allreduce(tensor, average=None, name=None, compression=Compression.none, op=None,
              prescale_factor=1.0, postscale_factor=1.0, process_set=None):  # pylint: disable=unused-argument
    """Average (default) or reduce ``tensor`` over all ranks; returns a new tensor.

    Differentiable: the gradient of an allreduce is the allreduce of the gradient."""
    comp, cctx = compression.compress(tensor)
    if comp.requires_grad and torch.is_grad_enabled():
        out = _AllreduceFn.apply(comp, average, name, op, prescale_factor, postscale_factor)
    else:
        out = synchronize(allreduce_async(comp, average, name, op, prescale_factor, postscale_factor))
    return compression.decompress(out, cctx)


def allreduce_(tensor, average=None, name=None, op=None, prescale_factor=1.0, postscale_factor=1.0,
               process_set=None):  # pylint: disable=unused-argument
    """In-place allreduce of ``tensor`` (blocking)."""
    return synchronize(allreduce_async_(tensor, average, name, op, prescale_factor, postscale_factor))


def grouped_allreduce_async(tensors, average=None, name=None, op=None, prescale_factor=1.0,
                            postscale_factor=1.0, process_set=None):  # pylint: disable=unused-argument
    """All tensors of the group are submitted together, so the engine fuses them into one
    launch whenever dtype/op agree and they fit the fusion threshold."""
    op = _resolve_op(average, op)
    base = _auto_name("grouped_allreduce", name)
    subs = []
    for i, t in enumerate(tensors):
        out = torch.empty_like(t, memory_format=torch.contiguous_format)
        subs.append(_allreduce_async_impl(t, out, "%s.%d" % (base, i), op, prescale_factor,
                                          postscale_factor))
    return _group(subs)


def grouped_allreduce_async_(tensors, average=None, name=None, op=None, prescale_factor=1.0,
                             postscale_factor=1.0, process_set=None):  # pylint: disable=unused-argument
    """In-place variant of ``grouped_allreduce_async``."""
    op = _resolve_op(average, op)
    base = _auto_name("grouped_allreduce", name)
    subs = [_allreduce_async_impl(t, t, "%s.%d" % (base, i), op, prescale_factor, postscale_factor)
            for i, t in enumerate(tensors)]
    return _group(subs)


def _group(subs):
    def post(hd, _subs=subs):
        return [synchronize(s) for s in _subs]
    return _Handle([], None, subs, post, False)


def grouped_allreduce(tensors, average=None, name=None, compression=Compression.none, op=None,
                      prescale_factor=1.0, postscale_factor=1.0, process_set=None):  # pylint: disable=unused-argument
    """Allreduce a list of tensors as one negotiated group (fused into as few launches as possible)."""
    pairs = [compression.compress(t) for t in tensors]
    outs = synchronize(grouped_allreduce_async([p[0] for p in pairs], average, name, op,
                                               prescale_factor, postscale_factor))
    return [compression.decompress(o, p[1]) for o, p in zip(outs, pairs)]


def grouped_allreduce_(tensors, average=None, name=None, op=None, prescale_factor=1.0,
                       postscale_factor=1.0, process_set=None):  # pylint: disable=unused-argument
    """In-place variant of ``grouped_allreduce``."""
    return synchronize(grouped_allreduce_async_(tensors, average, name, op, prescale_factor,
                                                postscale_factor))


# ---------------------------------------------------------------------------------------------
# allgather
# ---------------------------------------------------------------------------------------------
def _host_allgather_ints(values, name):
    """Small host-side allgather of a python int list (sizes, splits)."""
    t = torch.tensor(list(values), dtype=torch.int64).reshape(1, -1)
    h = _enqueue(_core.ALLGATHER, name, t, None)
    try:
        _S.engine.wait(h)
        out = torch.empty((_S.engine.out_dim0(h), t.shape[1]), dtype=torch.int64)
        _S.engine.copy_out(h, out.data_ptr())
    except _core.EngineError as e:
        raise HorovodInternalError(str(e))
    finally:
        _S.engine.release(h)
    return out


def allgather_async(tensor, name=None, process_set=None):  # pylint: disable=unused-argument
    """Concatenate ``tensor`` from all ranks along dim 0 (dim 0 may differ per rank)."""
    _require()
    name = _auto_name("allgather", name)
    src = tensor.contiguous()
    if src.dim() == 0:
        src = src.reshape(1)
    if src.is_cuda:
        # optimistic: size the output for equal shards and let the negotiation response (which carries every rank's
        # rows) confirm it - one round trip and no host sync.  Unequal shards: the engine leaves the buffers alone on
        # every rank and the exact-size path below runs at synchronize() time.
        n = size()
        out = torch.empty((n * src.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        h = _enqueue(_core.ALLGATHER, name, src, out, root=1)

        def exact(hd, _src=src, _h=h, _out=out, _name=name):
            if _S.engine.out_dim0(_h) == _out.shape[0]:
                return _out
            sizes = _host_allgather_ints([_src.shape[0]], _name + ".dim0").reshape(-1).tolist()
            full = torch.empty((sum(sizes),) + tuple(_src.shape[1:]), dtype=_src.dtype, device=_src.device)
            h2 = _enqueue(_core.ALLGATHER, _name + ".exact", _src, full)
            try:
                _S.engine.wait(h2)
                ev = _S.engine.done_event(h2)
                if ev:
                    _S.comm.current_stream_wait(ev, release=True)
            except _core.EngineError as e:
                raise HorovodInternalError(str(e))
            finally:
                _S.engine.release(h2)
            return full
        return _Handle([h], out, (src, out), exact, True)
    h = _enqueue(_core.ALLGATHER, name, src, None)

    def post(hd, _src=src, _h=h):
        out = torch.empty((_S.engine.out_dim0(_h),) + tuple(_src.shape[1:]), dtype=_src.dtype)
        _S.engine.copy_out(_h, out.data_ptr())
        return out
    return _Handle([h], None, (src,), post, False)


class _AllgatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, name):
        ctx.dim0 = tensor.shape[0] if tensor.dim() > 0 else 1
        sizes = _host_allgather_ints([ctx.dim0], _auto_name("allgather.bwd_sizes", None))
        ctx.offset = int(sizes.reshape(-1)[:rank()].sum())
        return synchronize(allgather_async(tensor, name))

    @staticmethod
    def backward(ctx, grad):
        g = allreduce(grad.contiguous(), op=Sum)
        return g.narrow(0, ctx.offset, ctx.dim0), None


def allgather(tensor, name=None, process_set=None):  # pylint: disable=unused-argument
    """Concatenate ``tensor`` of every rank along dim 0 (first dimensions may differ); differentiable."""
    if tensor.requires_grad and torch.is_grad_enabled():
        return _AllgatherFn.apply(tensor, name)
    return synchronize(allgather_async(tensor, name))


def grouped_allgather_async(tensors, name=None, process_set=None):  # pylint: disable=unused-argument
    """Start the allgather of every tensor in the list; one handle for the whole group."""
    base = _auto_name("grouped_allgather", name)
    return _group([allgather_async(t, "%s.%d" % (base, i)) for i, t in enumerate(tensors)])


def grouped_allgather(tensors, name=None, process_set=None):
    """Allgather every tensor in the list; returns the list of gathered tensors."""
    return synchronize(grouped_allgather_async(tensors, name, process_set))


# ---------------------------------------------------------------------------------------------
# broadcast
# ---------------------------------------------------------------------------------------------
def _broadcast_async_impl(tensor, output, root_rank, name):
    _require()
    if not 0 <= root_rank < size():
        raise ValueError("root_rank %d out of range for size %d" % (root_rank, size()))
    src = tensor if tensor.is_contiguous() else tensor.contiguous()
    dst = output if output.is_contiguous() else torch.empty_like(src)
    if dst is not src and rank() == root_rank:
        dst.copy_(src)
    h = _enqueue(_core.BROADCAST, name, dst, dst, root=root_rank)

    def post(hd, _dst=dst, _out=output):
        if _dst is not _out:
            _out.copy_(_dst)
        return _out
    return _Handle([h], output, (src, dst, tensor), post, tensor.is_cuda)


def broadcast_async(tensor, root_rank, name=None, process_set=None):  # pylint: disable=unused-argument
    """Start an out-of-place broadcast from ``root_rank``; returns a handle."""
    output = torch.empty_like(tensor, memory_format=torch.contiguous_format)
    return _broadcast_async_impl(tensor, output, root_rank, _auto_name("broadcast", name))


def broadcast_async_(tensor, root_rank, name=None, process_set=None):  # pylint: disable=unused-argument
    """Start an in-place broadcast from ``root_rank``; returns a handle."""
    return _broadcast_async_impl(tensor, tensor, root_rank, _auto_name("broadcast", name))


class _BroadcastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, root_rank, name):
        ctx.root = root_rank
        return synchronize(broadcast_async(tensor, root_rank, name))

    @staticmethod
    def backward(ctx, grad):
        g = allreduce(grad.contiguous(), op=Sum)
        if rank() != ctx.root:
            g = g * 0
        return g, None, None


def broadcast(tensor, root_rank, name=None, process_set=None):  # pylint: disable=unused-argument
    """Return ``root_rank``'s value of ``tensor`` on every rank; differentiable."""
    if tensor.requires_grad and torch.is_grad_enabled():
        return _BroadcastFn.apply(tensor, root_rank, name)
    return synchronize(broadcast_async(tensor, root_rank, name))


def broadcast_(tensor, root_rank, name=None, process_set=None):  # pylint: disable=unused-argument
    """Overwrite ``tensor`` with ``root_rank``'s value (blocking)."""
    return synchronize(broadcast_async_(tensor, root_rank, name))


# ---------------------------------------------------------------------------------------------
# alltoall / reducescatter
# ---------------------------------------------------------------------------------------------
def alltoall_async(tensor, splits=None, name=None, process_set=None):  # pylint: disable=unused-argument
    """Start an alltoall; ``splits[i]`` rows go to rank i (even split when omitted).  Returns a handle."""
    _require()
    name = _auto_name("alltoall", name)
    src = tensor.contiguous()
    if src.dim() == 0:
        raise ValueError("alltoall needs a tensor with at least one dimension")
    want_splits = splits is not None
    if splits is None:
        if src.shape[0] % size() != 0:
            raise ValueError("tensor first dimension must be divisible by size() when splits is None")
        sp = [src.shape[0] // size()] * size()
    else:
        sp = [int(s) for s in (splits.tolist() if torch.is_tensor(splits) else splits)]
        if len(sp) != size() or sum(sp) != src.shape[0]:
            raise ValueError("splits must have size() entries that sum to tensor.shape[0]")
    if src.is_cuda:
        matrix = _host_allgather_ints(sp, name + ".splits")            # [src, dst]
        recv = matrix[:, rank()].tolist()
        out = torch.empty((sum(recv),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        h = _enqueue(_core.ALLTOALL, name, src, out, splits=sp)
        rs = torch.tensor(recv, dtype=torch.int32)

        def post_c(hd, _out=out, _rs=rs):
            return (_out, _rs) if want_splits else _out
        return _Handle([h], out, (src, out), post_c, True)
    h = _enqueue(_core.ALLTOALL, name, src, None, splits=sp)

    def post(hd, _src=src, _h=h):
        out = torch.empty((_S.engine.out_dim0(_h),) + tuple(_src.shape[1:]), dtype=_src.dtype)
        _S.engine.copy_out(_h, out.data_ptr())
        if want_splits:
            return out, torch.tensor(_S.engine.recv_splits(_h), dtype=torch.int32)
        return out
    return _Handle([h], None, (src,), post, False)


def alltoall(tensor, splits=None, name=None, process_set=None):  # pylint: disable=unused-argument
    """Rank r sends ``splits[d]`` rows to rank d and receives the rows addressed to it."""
    return synchronize(alltoall_async(tensor, splits, name))


def reducescatter_async(tensor, name=None, op=Average, prescale_factor=1.0, postscale_factor=1.0,
                        process_set=None):  # pylint: disable=unused-argument
    """Start a reduce-scatter along dim 0 (rank r keeps the r-th slice of the reduction); returns a handle."""
    _require()
    name = _auto_name("reducescatter", name)
    if op == Average:
        op, postscale_factor = Sum, postscale_factor / float(size())
    src = tensor.contiguous()
    if src.dim() == 0:
        raise ValueError("reducescatter needs a tensor with at least one dimension")
    d0 = src.shape[0]
    rows = d0 // size() + (1 if rank() < d0 % size() else 0)
    if src.is_cuda:
        out = torch.empty((rows,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        h = _enqueue(_core.REDUCESCATTER, name, src, out, op=op, pre=prescale_factor,
                     post=postscale_factor)
        return _Handle([h], out, (src, out), None, True)
    h = _enqueue(_core.REDUCESCATTER, name, src, None, op=op, pre=prescale_factor,
                 post=postscale_factor)

    def post(hd, _src=src, _h=h, _rows=rows):
        out = torch.empty((_rows,) + tuple(_src.shape[1:]), dtype=_src.dtype)
        _S.engine.copy_out(_h, out.data_ptr())
        return out
    return _Handle([h], None, (src,), post, False)


def reducescatter(tensor, name=None, op=Average, prescale_factor=1.0, postscale_factor=1.0,
                  process_set=None):  # pylint: disable=unused-argument
    """Reduce over ranks, then rank r keeps the r-th slice of dim 0 (remainder rows go to the
    lowest ranks)."""
    return synchronize(reducescatter_async(tensor, name, op, prescale_factor, postscale_factor))


def grouped_reducescatter_async(tensors, name=None, op=Average, process_set=None):  # pylint: disable=unused-argument
    """Start the reduce-scatter of every tensor in the list; one handle for the whole group."""
    base = _auto_name("grouped_reducescatter", name)
    return _group([reducescatter_async(t, "%s.%d" % (base, i), op) for i, t in enumerate(tensors)])


def grouped_reducescatter(tensors, name=None, op=Average, process_set=None):
    """Reduce-scatter every tensor in the list; returns the list of local slices."""
    return synchronize(grouped_reducescatter_async(tensors, name, op, process_set))


# ---------------------------------------------------------------------------------------------
# barrier / join
# ---------------------------------------------------------------------------------------------
def barrier(process_set=None):  # pylint: disable=unused-argument
    """Block until every rank has reached the barrier."""
    _require()
    name = _auto_name("barrier", None)
    h = _S.engine.enqueue(_core.BARRIER, name, _core.U8, -1, 0, 0, [], 1.0, 1.0, 0, 0)
    try:
        _S.engine.wait(h)
    except _core.EngineError as e:
        raise HorovodInternalError(str(e))
    finally:
        _S.engine.release(h)


def join(device=-1):  # pylint: disable=unused-argument
    """Signal that this rank has run out of data.  Blocks until every rank has joined; while
    waiting, this rank contributes zeros to the allreduces of the others.  Returns the rank
    that joined last."""
    _require()
    h = _S.engine.enqueue(_core.JOIN, "__join__", _core.U8, -1, 0, 0, [], 1.0, 1.0, 0, 0)
    try:
        _S.engine.wait(h)
        return _S.engine.last_joined(h)
    except _core.EngineError as e:
        raise HorovodInternalError(str(e))
    finally:
        _S.engine.release(h)


# ---------------------------------------------------------------------------------------------
# objects, parameters, optimizer state
# ---------------------------------------------------------------------------------------------
def broadcast_object(obj, root_rank=0, name=None, process_set=None):  # pylint: disable=unused-argument
    """cloudpickle ``obj`` on ``root_rank`` and return it on every rank."""
    _require()
    import cloudpickle
    name = _auto_name("broadcast_object", name)
    if rank() == root_rank:
        buf = io.BytesIO()
        cloudpickle.dump(obj, buf)
        data = torch.frombuffer(bytearray(buf.getvalue()), dtype=torch.uint8)
        sz = torch.tensor([data.numel()], dtype=torch.int64)
    else:
        data = None
        sz = torch.zeros(1, dtype=torch.int64)
    broadcast_(sz, root_rank, name + ".sz")
    if rank() != root_rank:
        data = torch.empty(int(sz.item()), dtype=torch.uint8)
    broadcast_(data, root_rank, name + ".data")
    return pickle.loads(data.numpy().tobytes())


def allgather_object(obj, name=None, process_set=None):  # pylint: disable=unused-argument
    """Gather one picklable object per rank into a list ordered by rank."""
    _require()
    import cloudpickle
    name = _auto_name("allgather_object", name)
    payload = cloudpickle.dumps(obj)
    data = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    sizes = _host_allgather_ints([data.numel()], name + ".sz").reshape(-1).tolist()
    gathered = synchronize(allgather_async(data, name + ".data"))
    out, off = [], 0
    raw = gathered.numpy().tobytes()
    for s in sizes:
        out.append(pickle.loads(raw[off:off + s]))
        off += s
    return out


def broadcast_parameters(params, root_rank=0):
    """Broadcast a ``state_dict()`` / ``named_parameters()`` from ``root_rank`` to all ranks.

    CUDA tensors are grouped by dtype, flattened into one buffer and sent with a single
    peer-memory broadcast kernel per dtype instead of one launch per
    tensor."""
    _require()
    if isinstance(params, dict):
        items = sorted(params.items())
    elif isinstance(params, (list, tuple)) or hasattr(params, "__iter__"):
        items = list(params)
        if items and not isinstance(items[0], tuple):
            raise ValueError("invalid params of type: %s" % type(params))
    else:
        raise ValueError("invalid params of type: %s" % type(params))
    tensors = [(n, p) for n, p in items if torch.is_tensor(p)]
    groups = {}
    handles = []
    for n, p in tensors:
        t = p.data if isinstance(p, torch.nn.Parameter) else p
        if t.is_cuda and t.numel() > 0:
            groups.setdefault((t.dtype, t.device), []).append(t)
        else:
            handles.append(broadcast_async_(t, root_rank, "param." + n))
    for i, ((dt, dev), ts) in enumerate(sorted(groups.items(), key=lambda kv: str(kv[0]))):
        flat = torch.cat([t.reshape(-1) for t in ts])
        broadcast_(flat, root_rank, "param_group.%d.%s" % (i, str(dt)))
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
    for h in handles:
        synchronize(h)


def broadcast_optimizer_state(optimizer, root_rank=0, model=None):  # pylint: disable=unused-argument
    """Make every rank's optimizer state (momentum buffers, step counts, hyper-parameters)
    equal to ``root_rank``'s.  Used when resuming from a checkpoint loaded on rank 0."""
    _require()
    if isinstance(optimizer, torch.optim.LBFGS):
        raise ValueError("cannot broadcast torch.optim.LBFGS state")
    if getattr(optimizer, "_fused", False) and getattr(optimizer, "_buckets", None):
        # The fused optimizer's moments / master weights are ONE sharded global copy (its state_dict() is a
        # collective), so only the plain part travels - plus the checkpoint payload the root loaded since the
        # last step, which gives the usual "load on rank 0, then broadcast" resume recipe its meaning.
        payload = None
        if rank() == root_rank:
            payload = (torch.optim.Optimizer.state_dict(optimizer), getattr(optimizer, "_pending_fused", None))
        plain, pending = broadcast_object(payload, root_rank, "optimizer_state")
        if rank() != root_rank:
            sd = dict(plain)
            if pending is not None:
                sd["sparkdl_fused"] = pending
            optimizer.load_state_dict(sd)
        return
    sd = optimizer.state_dict() if rank() == root_rank else None
    sd = broadcast_object(sd, root_rank, "optimizer_state")
    if rank() != root_rank:
        optimizer.load_state_dict(sd)


from spark_deep_learning_b200.parallel.optimizer import DistributedOptimizer  # noqa: E402,F401
from spark_deep_learning_b200.parallel.sync_batch_norm import SyncBatchNorm  # noqa: E402,F401


from spark_deep_learning_b200.parallel import elastic  # noqa: E402  pylint: disable=wrong-import-position
