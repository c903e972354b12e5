"""ctypes binding of ``libsparkdl_core.so`` (C++ engine: negotiation, fusion, shm collectives).

The library is built in-tree by ``__graft_entry__.build()`` / :func:`spark_deep_learning_b200.build.build_core`.
It fails loudly if missing: there is no pure-Python fallback for the engine.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

# dtype codes shared with csrc/core/common.h and csrc/cuda/*.cuh
U8, I8, I32, I64, F16, BF16, F32, F64, BOOL, F8E4M3 = range(10)
# request types
ALLREDUCE, ALLGATHER, BROADCAST, ALLTOALL, REDUCESCATTER, BARRIER, JOIN, SHUTDOWN = range(8)
# reduce ops
OP_SUM, OP_AVERAGE, OP_MIN, OP_MAX, OP_PRODUCT = range(5)


def lib_path():
    """The engine library; ``SPARKDL_CORE_LIB`` points at an alternative build (e.g. one compiled with
    ``-fsanitize=address`` for a sanitizer run of the Python test-suite)."""
    return os.environ.get("SPARKDL_CORE_LIB") or os.path.join(_HERE, "libsparkdl_core.so")


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        from spark_deep_learning_b200 import build as _build
        _build.build_core()
    lib = ctypes.CDLL(path)
    vp, i64, u64, i32, u32, dbl, cp = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64,
                                        ctypes.c_int, ctypes.c_uint32, ctypes.c_double,
                                        ctypes.c_char_p)
    lib.sdl_engine_create.restype = vp
    lib.sdl_engine_init.restype = cp
    lib.sdl_engine_init.argtypes = [vp, cp, i32, i32, u64, i64, i32, dbl, dbl, dbl, cp]
    lib.sdl_engine_shutdown.argtypes = [vp]
    lib.sdl_engine_destroy.argtypes = [vp]
    lib.sdl_engine_abort.argtypes = [vp]
    lib.sdl_set_gpu_executor.argtypes = [vp, vp, vp]
    lib.sdl_enqueue.restype = i64
    lib.sdl_enqueue.argtypes = [vp, u32, cp, u32, i32, u32, i32, i32, ctypes.POINTER(i64), dbl, dbl,
                                vp, vp, ctypes.POINTER(ctypes.c_int32), u64]
    lib.sdl_poll.argtypes = [vp, i64]
    lib.sdl_wait.argtypes = [vp, i64, dbl]
    lib.sdl_handle_error.argtypes = [vp, i64, cp, i32]
    lib.sdl_handle_out_bytes.restype = i64
    lib.sdl_handle_out_bytes.argtypes = [vp, i64]
    lib.sdl_handle_out_dim0.restype = i64
    lib.sdl_handle_out_dim0.argtypes = [vp, i64]
    lib.sdl_handle_copy_out.argtypes = [vp, i64, vp]
    lib.sdl_handle_recv_splits.argtypes = [vp, i64, ctypes.POINTER(ctypes.c_int32)]
    lib.sdl_handle_last_joined.argtypes = [vp, i64]
    lib.sdl_handle_done_event.restype = u64
    lib.sdl_handle_done_event.argtypes = [vp, i64]
    lib.sdl_release.argtypes = [vp, i64]
    lib.sdl_stat_responses.restype = u64
    lib.sdl_stat_responses.argtypes = [vp]
    lib.sdl_stat_tensors.restype = u64
    lib.sdl_stat_tensors.argtypes = [vp]
    lib.sdl_reduce_into.argtypes = [vp, vp, i64, u32, u32]
    lib.sdl_scale_buffer.argtypes = [vp, i64, u32, dbl]
    if hasattr(lib, "sdl_timeline_start"):
        lib.sdl_timeline_start.argtypes = [vp, cp]
        lib.sdl_timeline_stop.argtypes = [vp]
    if hasattr(lib, "sdl_gbdt_hist_cpu"):          # absent only in a stale in-tree build
        lib.sdl_gbdt_hist_cpu.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32]
    if hasattr(lib, "sdl_stat_cache_hits"):
        for fn in ("sdl_stat_cache_hits", "sdl_stat_cache_misses", "sdl_stat_cache_size"):
            getattr(lib, fn).restype = u64
            getattr(lib, fn).argtypes = [vp]
        lib.sdl_stat_negotiation_us.restype = dbl
        lib.sdl_stat_negotiation_us.argtypes = [vp]
        lib.sdl_set_cache_enabled.argtypes = [vp, i32]
    if hasattr(lib, "sdl_exit_watchdog"):
        lib.sdl_exit_watchdog.argtypes = [dbl, i32]
        lib.sdl_exit_watchdog.restype = None
    _LIB = lib
    return lib


def exit_watchdog(seconds, code):
    """Native detached thread that ``_exit(code)``s after ``seconds`` unless the process has left by itself."""
    lib = load()
    if hasattr(lib, "sdl_stat_cache_hits"):
        for fn in ("sdl_stat_cache_hits", "sdl_stat_cache_misses", "sdl_stat_cache_size"):
            getattr(lib, fn).restype = u64
            getattr(lib, fn).argtypes = [vp]
        lib.sdl_stat_negotiation_us.restype = dbl
        lib.sdl_stat_negotiation_us.argtypes = [vp]
        lib.sdl_set_cache_enabled.argtypes = [vp, i32]
    if hasattr(lib, "sdl_exit_watchdog"):
        lib.sdl_exit_watchdog(float(seconds), int(code))


class EngineError(RuntimeError):
    pass


class CoreEngine(object):
    """One per process.  Thin object wrapper over the C ABI."""

    def __init__(self, shm_name, rank, size, cfg):
        self._lib = load()
        self._e = ctypes.c_void_p(self._lib.sdl_engine_create())
        self.rank, self.size = rank, size
        err = self._lib.sdl_engine_init(
            self._e, shm_name.encode(), rank, size, int(cfg.shm_stage_bytes),
            int(cfg.fusion_threshold), max(50, int(cfg.cycle_time_ms * 1000)),
            float(cfg.stall_check_s), float(cfg.stall_shutdown_s), float(cfg.start_timeout_s),
            cfg.timeline.encode())
        if err:
            raise EngineError("sparkdl engine init failed: %s" % err.decode())
        self._alive = True

    def enqueue(self, req_type, name, dtype, device, op, root, shape, prescale, postscale,
                in_ptr, out_ptr, splits=None, ready_event=0):
        nd = len(shape)
        shp = (ctypes.c_int64 * max(nd, 1))(*shape)
        sp = None
        if splits is not None:
            sp = (ctypes.c_int32 * len(splits))(*[int(s) for s in splits])
        return self._lib.sdl_enqueue(self._e, req_type, name.encode()[:127], dtype, device, op, root,
                                     nd, shp, float(prescale), float(postscale),
                                     ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr), sp,
                                     int(ready_event))

    def poll(self, h):
        return self._lib.sdl_poll(self._e, h)

    def wait(self, h, timeout=None):
        """Block until the handle completes.  Raises :class:`EngineError` on failure."""
        while True:
            st = self._lib.sdl_wait(self._e, h, 1.0 if timeout is None else float(timeout))
            if st == 1:
                return
            if st < 0:
                raise EngineError(self.error(h))
            if timeout is not None:
                raise TimeoutError("collective did not complete within %.1fs" % timeout)

    def error(self, h):
        buf = ctypes.create_string_buffer(1024)
        self._lib.sdl_handle_error(self._e, h, buf, 1024)
        return buf.value.decode("utf-8", "replace")

    def out_bytes(self, h):
        return self._lib.sdl_handle_out_bytes(self._e, h)

    def out_dim0(self, h):
        return self._lib.sdl_handle_out_dim0(self._e, h)

    def copy_out(self, h, ptr):
        self._lib.sdl_handle_copy_out(self._e, h, ctypes.c_void_p(ptr))

    def recv_splits(self, h):
        arr = (ctypes.c_int32 * self.size)()
        n = self._lib.sdl_handle_recv_splits(self._e, h, arr)
        return list(arr[:max(n, 0)])

    def last_joined(self, h):
        return self._lib.sdl_handle_last_joined(self._e, h)

    def done_event(self, h):
        return self._lib.sdl_handle_done_event(self._e, h)

    def release(self, h):
        self._lib.sdl_release(self._e, h)

    def set_gpu_executor(self, fn_addr, ctx_addr):
        self._lib.sdl_set_gpu_executor(self._e, ctypes.c_void_p(fn_addr), ctypes.c_void_p(ctx_addr))

    def stats(self):
        out = {"responses": self._lib.sdl_stat_responses(self._e),
               "tensors": self._lib.sdl_stat_tensors(self._e)}
        if hasattr(self._lib, "sdl_stat_cache_hits"):
            out.update({"cache_hits": self._lib.sdl_stat_cache_hits(self._e),
                        "cache_misses": self._lib.sdl_stat_cache_misses(self._e),
                        "cache_size": self._lib.sdl_stat_cache_size(self._e),
                        "negotiation_us": self._lib.sdl_stat_negotiation_us(self._e)})
        return out

    def set_cache_enabled(self, on):
        if hasattr(self._lib, "sdl_set_cache_enabled"):
            self._lib.sdl_set_cache_enabled(self._e, 1 if on else 0)

    def timeline_start(self, path):
        return self._lib.sdl_timeline_start(self._e, path.encode()) == 0

    def timeline_stop(self):
        self._lib.sdl_timeline_stop(self._e)

    def abort(self):
        if self._alive:
            self._lib.sdl_engine_abort(self._e)

    def shutdown(self):
        if self._alive:
            self._alive = False
            self._lib.sdl_engine_shutdown(self._e)
            self._lib.sdl_engine_destroy(self._e)
