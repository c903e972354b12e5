"""Symmetric-memory communicator: Python orchestration of ``libsparkdl_cuda.so``.

What BASELINE.json's north star asks for at ``hvd.init()``: "exchanges cudaIpc/fabric handles
and NVLink topology at init".  Concretely:

* probe: peer-access matrix, VMM posix-fd export, NVSwitch multicast capability;
* allocate per-rank symmetric buffers with CUDA VMM, pass the file descriptors to the peers over
  unix sockets (``SCM_RIGHTS``), map every peer's buffer locally; fall back to ``cudaIpc``;
* bind each buffer to an NVSwitch multicast object so kernels can use ``multimem.*`` (NVLS);
* signal pads (flag words in peer memory) for in-kernel barriers, one set per channel.

Channels: 0 = engine (generic hvd.* calls), 1 = DistributedOptimizer buckets, 2 = GBDT, 3 = spare.
"""
import ctypes
import os

import torch

from spark_deep_learning_b200.parallel import core as _core
from spark_deep_learning_b200.parallel.store import FdServer, fetch_fd, store_barrier
from spark_deep_learning_b200.utils import nvtx

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

WIRE = {None: 0, "bf16": 1, "fp16": 2, "fp8": 3}
ALGO = {"auto": 0, "oneshot": 1, "twoshot": 2, "multimem": 3, "ll": 4}
_TORCH_DTYPE = {
    torch.uint8: _core.U8, torch.int8: _core.I8, torch.int32: _core.I32, torch.int64: _core.I64,
    torch.float16: _core.F16, torch.bfloat16: _core.BF16, torch.float32: _core.F32,
    torch.float64: _core.F64, torch.bool: _core.BOOL,
}


def lib_path():
    return os.path.join(_PKG, "libsparkdl_cuda.so")


def load():
    """Load the CUDA library; build it first if the .so is missing.  Never falls back silently."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(lib_path()):
        from spark_deep_learning_b200 import build as _build
        _build.build_cuda()
    lib = ctypes.CDLL(lib_path())
    vp, i32, i64, u64, f32, dbl, sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong,
                                        ctypes.c_ulonglong, ctypes.c_float, ctypes.c_double,
                                        ctypes.c_size_t)
    P = ctypes.POINTER
    sig = {
        "sdlc_comm_create": (vp, [i32, i32, i32, dbl, i32, i64, i32]),
        "sdlc_comm_destroy": (None, [vp]),
        "sdlc_last_error": (ctypes.c_char_p, [vp]),
        "sdlc_probe": (i32, [vp, P(i32)]),
        "sdlc_buf_alloc": (i32, [vp, sz, i32]),
        "sdlc_buf_export_fd": (i32, [vp, i32]),
        "sdlc_buf_import_fd": (i32, [vp, i32, i32, i32]),
        "sdlc_buf_export_ipc": (i32, [vp, i32, ctypes.c_char_p]),
        "sdlc_buf_import_ipc": (i32, [vp, i32, i32, ctypes.c_char_p]),
        "sdlc_buf_set_peer_ptr": (i32, [vp, i32, i32, vp]),
        "sdlc_buf_ptr": (vp, [vp, i32, i32]),
        "sdlc_buf_mc_ptr": (vp, [vp, i32]),
        "sdlc_buf_bytes": (sz, [vp, i32]),
        "sdlc_buf_alloc_bytes": (sz, [vp, i32]),
        "sdlc_mc_create": (i32, [vp, i32]),
        "sdlc_mc_import": (i32, [vp, i32, i32]),
        "sdlc_mc_add_device": (i32, [vp, i32]),
        "sdlc_mc_bind": (i32, [vp, i32]),
        "sdlc_comm_set_pads": (i32, [vp, i32]),
        "sdlc_pad_bytes": (sz, []),
        "sdlc_comm_set_staging": (i32, [vp, i32, i32]),
        "sdlc_comm_set_use_mc": (None, [vp, i32]),
        "sdlc_comm_set_max_blocks": (None, [vp, i32]),
        "sdlc_comm_set_oneshot_max": (None, [vp, i64]),
        "sdlc_channel_stream": (u64, [vp, i32]),
        "sdlc_record_event": (u64, [vp, u64]),
        "sdlc_stream_wait_event": (i32, [u64, u64]),
        "sdlc_event_synchronize": (i32, [u64]),
        "sdlc_release_event": (None, [vp, u64]),
        "sdlc_event_pool_size": (u64, [vp]),
        "sdlc_check_error": (ctypes.c_uint, [vp]),
        "sdlc_clear_error": (None, [vp]),
        "sdlc_launch_count": (u64, [vp]),
        "sdlc_comm_set_ll": (i32, [vp, i32]),
        "sdlc_ll_bytes": (ctypes.c_size_t, []),
        "sdlc_max_peers": (i32, []),
        "sdlc_allreduce": (i32, [vp, i32, i32, i64, i64, i32, i32, f32, f32, i32, u64]),
        "sdlc_wire_bytes": (i64, [i64, i32]),
        "sdlc_allreduce_wire": (i32, [vp, i32, vp, i64, i32, i32, i32, i64, f32, f32, u64]),
        "sdlc_barrier": (i32, [vp, i32, u64]),
        "sdlc_pull": (i32, [vp, i32, i32, vp, P(i64), P(i64), P(i64), i64, u64]),
        "sdlc_bcast_multimem": (i32, [vp, i32, i32, vp, vp, i64, i32, u64]),
        "sdlc_reducescatter": (i32, [vp, i32, i32, vp, i64, i64, i32, i32, f32, f32, i64, u64]),
        "sdlc_pack": (i32, [vp, i32, P(u64), P(i64), P(i64), i32, i32, u64]),
        "sdlc_scale": (i32, [vp, vp, i64, i32, f32, u64]),
        "sdlc_fused_step": (i32, [vp, i32, i32, i32, i64, i32, vp, vp, vp, f32, vp, u64]),
        "sdlc_step_advance": (i32, [vp, i32, vp, u64]),
        "sdlc_fused_state_elems": (i64, [vp, i64, i32]),
        "sdlc_fused_slice_start": (i64, [vp, i64, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


class CommError(RuntimeError):
    pass


class SymmetricBuffer(object):
    """A buffer that exists on every rank and is mapped into every rank's address space."""

    def __init__(self, comm, buf_id, nbytes):
        self.comm = comm
        self.id = buf_id
        self.nbytes = nbytes
        self._base = None

    def ptr(self, peer=None):
        peer = self.comm.rank if peer is None else peer
        return self.comm._lib.sdlc_buf_ptr(self.comm._h, self.id, peer)

    @property
    def multicast(self):
        return bool(self.comm._lib.sdlc_buf_mc_ptr(self.comm._h, self.id))

    def tensor(self, dtype, numel, offset_bytes=0):
        """A torch tensor aliasing this rank's copy (no ownership: the communicator owns the memory)."""
        if self._base is None:
            self._base = _tensor_from_ptr(self.ptr(), self.nbytes, self.comm.device)
        esz = torch.empty(0, dtype=dtype).element_size()
        return self._base[offset_bytes:offset_bytes + numel * esz].view(dtype)


def _tensor_from_ptr(ptr, nbytes, device):
    """Wrap raw device memory as a uint8 torch tensor through ``__cuda_array_interface__``."""
    class _Holder(object):
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                  "version": 3, "strides": None}
    t = torch.as_tensor(h, device=torch.device("cuda", device))
    t._sparkdl_holder = h  # keep alive
    return t


class SymmComm(object):
    """One per process (and per virtual rank in single-process tests)."""

    def __init__(self, ctx, cfg, device=None, exchange=True):
        self._lib = load()
        self.rank, self.world = ctx.rank, ctx.size
        self.ctx, self.cfg = ctx, cfg
        if self.world > 8:
            raise CommError("the symmetric-memory communicator supports up to 8 GPUs on one box, got world=%d "
                            "(docs/LIMITS.md)" % self.world)
        self.device = torch.cuda.current_device() if device is None else device
        self._h = ctypes.c_void_p(self._lib.sdlc_comm_create(
            self.rank, self.world, self.device, float(cfg.kernel_timeout_ms), int(cfg.comm_ctas),
            int(cfg.oneshot_max_bytes), 1 if cfg.use_multicast else 0))
        if not self._h:
            raise CommError("sdlc_comm_create failed")
        self._buffers = []
        self._fd_server = None
        self._n_alloc = 0
        self.transport = "local"
        self.multicast = False
        self.caps = self.probe()
        self._wire_bufs = {}
        self._opt_state = {}
        if exchange:
            self._bootstrap()

    # ---------------------------------------------------------------------------------------------
    def _err(self):
        return self._lib.sdlc_last_error(self._h).decode("utf-8", "replace")

    def probe(self):
        out = (ctypes.c_int * 8)()
        self._lib.sdlc_probe(self._h, out)
        return {"vmm_fd": bool(out[0]), "multicast": bool(out[1]), "devices": out[2],
                "peer_mask": out[3], "sms": out[4]}

    def _bootstrap(self):
        ctx, cfg = self.ctx, self.cfg
        store = ctx.store
        # 1. agree on transport + multicast: every rank publishes its capabilities
        store.set("comm/caps/%d" % self.rank, dict(self.caps, device=self.device))
        caps = [store.get("comm/caps/%d" % r) for r in range(self.world)]
        self.topology = caps
        want = cfg.transport
        all_vmm = all(c["vmm_fd"] for c in caps)
        if self.world == 1:
            self.transport = "local"
        elif want == "ipc" or not all_vmm:
            self.transport = "ipc"
        else:
            self.transport = "vmm"
        same_device = len({c["device"] for c in caps}) < self.world
        self.multicast = (self.transport == "vmm" and cfg.use_multicast and not same_device
                          and all(c["multicast"] for c in caps))
        if self.world > 1 and not same_device:
            need = 0
            for c in caps:
                need |= 1 << c["device"]
            if (self.caps["peer_mask"] & need) != need:
                raise CommError("GPU %d cannot access all peers (mask %x, need %x)"
                                % (self.device, self.caps["peer_mask"], need))
        if self.transport == "vmm":
            self._fd_server = FdServer(os.path.join(ctx.job_dir, "fd-%s-%d.sock" % (ctx.job_id[:24], self.rank)))
            store.set("comm/fdsock/%d" % self.rank, self._fd_server.path)
            self._fd_paths = [store.get("comm/fdsock/%d" % r) for r in range(self.world)]
        # 2. signal pads + per-channel staging buffers
        try:
            pads = self.alloc_symmetric(self._lib.sdlc_pad_bytes(), multicast=False)
        except CommError:
            if self.transport != "vmm":
                raise
            # VMM export/import failed on this box: fall back to cudaIpc for everything
            self._vote_transport_fallback()
            pads = self.alloc_symmetric(self._lib.sdlc_pad_bytes(), multicast=False)
        if self._lib.sdlc_comm_set_pads(self._h, pads.id):
            raise CommError(self._err())
        self.pads = pads
        if self.world > 1 and os.environ.get("SPARKDL_LL_ALLREDUCE", "1") != "0":
            # receive slots of the barrier-free small-message allreduce (<= 64 KB), zero-filled by the allocator
            self._ll = self.alloc_symmetric(self._lib.sdlc_ll_bytes(), multicast=False)
            if self._lib.sdlc_comm_set_ll(self._h, self._ll.id):
                raise CommError(self._err())
        staging_bytes = max(int(cfg.fusion_threshold) + (4 << 20), 16 << 20)
        self.staging = self.alloc_symmetric(staging_bytes)
        self._lib.sdlc_comm_set_staging(self._h, 0, self.staging.id)
        self.sync_all()
        if cfg.autotune and self.world > 1:
            self.autotune()

    def _vote_transport_fallback(self):
        self.transport = "ipc"
        self.multicast = False

    # ---------------------------------------------------------------------------------------------
    def _vote(self, tag, err):
        """Publish this rank's verdict (empty string = ok) and return the first error of any rank."""
        if self.world == 1:
            return err
        store = self.ctx.store
        store.set("comm/vote/%s/%d" % (tag, self.rank), err or "")
        errs = [store.get("comm/vote/%s/%d" % (tag, r)) for r in range(self.world)]
        bad = [e for e in errs if e]
        return bad[0] if bad else ""

    def alloc_symmetric(self, nbytes, multicast=None):
        """Collective: every rank allocates ``nbytes`` and maps every peer's allocation.

        Each phase (allocate+export, import) ends with a vote, so a failure on one rank raises
        :class:`CommError` on *every* rank instead of leaving the others waiting."""
        lib, store = self._lib, self.ctx.store if self.ctx is not None else None
        nbytes = int((nbytes + 255) // 256 * 256)
        seq = self._n_alloc
        self._n_alloc += 1
        kind = 1 if self.transport == "vmm" else 0
        # phase 1: allocate locally and export a shareable descriptor
        bid = lib.sdlc_buf_alloc(self._h, nbytes, kind)
        err = "" if bid >= 0 else "alloc: " + self._err()
        if not err and self.world > 1:
            if self.transport == "vmm":
                fd = lib.sdlc_buf_export_fd(self._h, bid)
                if fd < 0:
                    err = "cuMemExportToShareableHandle: " + self._err()
                self._fd_server.publish("buf/%d" % seq, fd if fd >= 0 else None)
            else:
                hb = ctypes.create_string_buffer(64)
                if lib.sdlc_buf_export_ipc(self._h, bid, hb):
                    err = "cudaIpcGetMemHandle: " + self._err()
                store.set("comm/ipc/%d/%d" % (seq, self.rank), (hb.raw, self.device))
        elif self.world > 1 and self.transport == "vmm":
            self._fd_server.publish("buf/%d" % seq, None)
        bad = self._vote("alloc%d/export" % seq, err)
        if bad:
            raise CommError("symmetric allocation failed (%s transport): %s" % (self.transport, bad))
        buf = SymmetricBuffer(self, bid, nbytes)
        # phase 2: map every peer's allocation
        if self.world > 1:
            err = ""
            try:
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    if self.transport == "vmm":
                        pfd = fetch_fd(self._fd_paths[r], "buf/%d" % seq)
                        if lib.sdlc_buf_import_fd(self._h, bid, r, pfd):
                            raise CommError("cuMemImportFromShareableHandle(rank %d): %s" % (r, self._err()))
                    else:
                        raw, _ = store.get("comm/ipc/%d/%d" % (seq, r))
                        if lib.sdlc_buf_import_ipc(self._h, bid, r, raw):
                            raise CommError("cudaIpcOpenMemHandle(rank %d): %s" % (r, self._err()))
            except Exception as e:  # pylint: disable=broad-except
                err = str(e)
            bad = self._vote("alloc%d/import" % seq, err)
            if bad:
                raise CommError("symmetric mapping failed (%s transport): %s" % (self.transport, bad))
            use_mc = self.multicast if multicast is None else (multicast and self.multicast)
            if use_mc:
                self._bind_multicast(bid, seq)
            store_barrier(store, "alloc/%d" % seq, self.rank, self.world)
        self._buffers.append(buf)
        return buf

    def _bind_multicast(self, bid, seq):
        lib, store = self._lib, self.ctx.store
        ok = True
        if self.rank == 0:
            fd = lib.sdlc_mc_create(self._h, bid)
            ok = fd >= 0
            if ok:
                self._fd_server.publish("mc/%d" % seq, fd)
            store.set("comm/mc/%d/created" % seq, ok)
        else:
            ok = store.get("comm/mc/%d/created" % seq)
            if ok:
                fd = fetch_fd(self._fd_paths[0], "mc/%d" % seq)
                ok = lib.sdlc_mc_import(self._h, bid, fd) == 0
        if ok:
            ok = lib.sdlc_mc_add_device(self._h, bid) == 0
        store.set("comm/mc/%d/added/%d" % (seq, self.rank), ok)
        all_ok = all(store.get("comm/mc/%d/added/%d" % (seq, r)) for r in range(self.world))
        if all_ok:
            ok = lib.sdlc_mc_bind(self._h, bid) == 0
            store.set("comm/mc/%d/bound/%d" % (seq, self.rank), ok)
            all_ok = all(store.get("comm/mc/%d/bound/%d" % (seq, r)) for r in range(self.world))
        if not all_ok:
            # NVLS is optional: run P2P-only from here on (every rank reaches the same verdict)
            self.multicast = False
            lib.sdlc_comm_set_use_mc(self._h, 0)

    # ---------------------------------------------------------------------------------------------
    def sync_all(self):
        """Device-wide sync + host barrier (init paths only)."""
        torch.cuda.synchronize(self.device)
        if self.world > 1 and self.ctx is not None:
            n = getattr(self, "_sync_seq", 0)
            self._sync_seq = n + 1
            store_barrier(self.ctx.store, "sync/%d" % n, self.rank, self.world)

    # ---------------------------------------------------------------------------------------------
    @nvtx.annotate("sparkdl.autotune")
    def autotune(self, sizes=(4 << 10, 16 << 10, 64 << 10, 256 << 10, 1 << 20), iters=20):
        """``SPARKDL_AUTOTUNE=1``: time one-shot vs two-shot/NVLS on this box for a few small
        sizes, move the one-shot threshold to the measured crossover, and cache the verdict per (world, transport)
        in ``~/.cache/sparkdl_b200/autotune.json`` so later jobs skip the measurement.  Collective: every rank
        runs the same launches; rank 0's timing decides (broadcast through the store)."""
        import json
        key = "w%d-%s-mc%d" % (self.world, self.transport, int(self.multicast))
        path = os.path.join(os.path.expanduser("~"), ".cache", "sparkdl_b200", "autotune.json")
        store = self.ctx.store
        cached = None
        if self.rank == 0 and os.path.exists(path):
            try:
                cached = json.load(open(path)).get(key)
            except (OSError, ValueError):
                cached = None
        if self.rank == 0:
            store.set("comm/autotune/cached", cached)
        cached = store.get("comm/autotune/cached")
        if cached is None:
            lib = self._lib
            s = self.channel_stream(3)
            ts = torch.cuda.ExternalStream(s, device=torch.device("cuda", self.device))
            best = 0
            big_algo = 3 if (self.multicast and self.staging.multicast) else 2
            for nb in sizes:
                t = {}
                for algo in (1, big_algo):
                    for _ in range(3):
                        lib.sdlc_allreduce(self._h, 3, self.staging.id, 0, nb // 4, _core.F32, 0, 1.0, 1.0, algo, s)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(ts)
                    for _ in range(iters):
                        lib.sdlc_allreduce(self._h, 3, self.staging.id, 0, nb // 4, _core.F32, 0, 1.0, 1.0, algo, s)
                    e1.record(ts)
                    ts.synchronize()
                    t[algo] = e0.elapsed_time(e1) / iters
                if t[1] <= t[big_algo]:
                    best = nb
            cached = {"oneshot_max_bytes": int(best)}
            if self.rank == 0:
                store.set("comm/autotune/result", cached)
                try:
                    os.makedirs(os.path.dirname(path), exist_ok=True)
                    allv = json.load(open(path)) if os.path.exists(path) else {}
                    allv[key] = cached
                    json.dump(allv, open(path, "w"), indent=1)
                except (OSError, ValueError):
                    pass
            cached = store.get("comm/autotune/result")
        self._lib.sdlc_comm_set_oneshot_max(self._h, int(cached["oneshot_max_bytes"]))
        self.autotuned = cached
        return cached

    def engine_executor(self):
        fn = ctypes.cast(self._lib.sdlc_engine_exec, ctypes.c_void_p).value
        return fn, self._h.value

    def channel_stream(self, channel):
        return self._lib.sdlc_channel_stream(self._h, channel)

    def record_ready_event(self):
        """Record an event on the current torch stream; returns the raw cudaEvent_t."""
        ev = self._lib.sdlc_record_event(self._h, torch.cuda.current_stream(self.device).cuda_stream)
        if not ev:
            raise CommError(self._err())
        return ev

    def current_stream_wait(self, event, release=False):
        """Make the current torch stream wait for ``event``; ``release=True`` hands the event back to the
        pool afterwards (the caller will not use it again)."""
        self._lib.sdlc_stream_wait_event(torch.cuda.current_stream(self.device).cuda_stream, event)
        if release:
            self._lib.sdlc_release_event(self._h, event)

    def release_event(self, event):
        if event:
            self._lib.sdlc_release_event(self._h, event)

    def _stream_wait_current(self, stream):
        """``stream`` waits for everything queued on the current torch stream so far."""
        ev = self.record_ready_event()
        self._lib.sdlc_stream_wait_event(stream, ev)
        self._lib.sdlc_release_event(self._h, ev)

    @property
    def event_pool_size(self):
        return self._lib.sdlc_event_pool_size(self._h)

    def check_error(self):
        e = self._lib.sdlc_check_error(self._h)
        if e:
            raise CommError("a collective kernel timed out waiting for a peer (code 0x%x: rank %d "
                            "waited for rank %d)" % (e, e & 0xff, (e >> 8) & 0xff))

    @property
    def launches(self):
        return self._lib.sdlc_launch_count(self._h)

    # ---------------------------------------------------------------------------------------------
    # collectives on symmetric buffers (hot path; no negotiation, callers guarantee ordering)
    # ---------------------------------------------------------------------------------------------
    @nvtx.annotate("sparkdl.allreduce")
    def allreduce_symmetric(self, buf, numel, dtype, op=_core.OP_SUM, pre=1.0, post=1.0, wire=None,
                            channel=1, algo="auto", offset_bytes=0, wait_current=True, stream=None):
        """In-place allreduce of a symmetric buffer on the channel's stream.  Returns the raw
        cudaEvent_t that marks completion (wait on it with :meth:`current_stream_wait`)."""
        lib = self._lib
        s = self.channel_stream(channel) if stream is None else stream
        if wait_current:
            self._stream_wait_current(s)
        code = _TORCH_DTYPE[dtype] if isinstance(dtype, torch.dtype) else int(dtype)
        if wire:
            if op not in (_core.OP_SUM, _core.OP_AVERAGE):
                raise CommError("wire formats support sum/average only")
            if numel % 8:
                raise CommError("wire-format allreduce needs a multiple of 8 elements, got %d (docs/LIMITS.md)" % numel)
            if op == _core.OP_AVERAGE:
                post = post / self.world
            wb = self._wire_buffer(buf, numel, wire)
            rc = lib.sdlc_allreduce_wire(self._h, channel, ctypes.c_void_p(buf.ptr() + offset_bytes), numel,
                                         code, WIRE[wire], wb.id, 0, pre, post, s)
        else:
            rc = lib.sdlc_allreduce(self._h, channel, buf.id, offset_bytes, numel, code, op, pre, post,
                                    ALGO[algo], s)
        if rc < 0:
            raise CommError("allreduce launch failed: %s" % self._err())
        ev = lib.sdlc_record_event(self._h, s)
        return ev

    def _wire_buffer(self, buf, numel, wire):
        key = (buf.id, wire)
        wb = self._wire_bufs.get(key)
        need = self._lib.sdlc_wire_bytes(numel, WIRE[wire])
        if wb is None or wb.nbytes < need:
            wb = self.alloc_symmetric(need, multicast=False)
            self._wire_bufs[key] = wb
        return wb

    @nvtx.annotate("sparkdl.barrier")
    def barrier(self, channel=0):
        if self._lib.sdlc_barrier(self._h, channel, 0):
            raise CommError(self._err())

    # ---------------------------------------------------------------------------------------------
    # fused reduce-scatter + optimizer + all-gather (DistributedOptimizer(fused=True))
    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def hyper_values(opt, group):
        """The 9 static entries of a param group's device hyper-parameter block (see fused_optim.cu)."""
        if isinstance(opt, torch.optim.SGD):
            return [0, group["lr"], group["weight_decay"], group["momentum"], group["dampening"],
                    1.0 if group["nesterov"] else 0.0, 0.0, 0.0, 0.0]
        b1, b2 = group["betas"]
        kind = 2 if isinstance(opt, torch.optim.AdamW) else 1
        return [kind, group["lr"], group["weight_decay"], 0.0, 0.0, 0.0, b1, b2, group["eps"]]

    @nvtx.annotate("sparkdl.fused_step")
    def fused_optimizer_step(self, opt, b, pre, post):
        """Launch the fused reduce-scatter + optimizer + all-gather kernel for bucket ``b``."""
        lib = self._lib
        s = self.channel_stream(1)
        code = _TORCH_DTYPE[b.dtype]
        st = b.state
        gi = opt._bucket_group[b.index]
        group = opt.param_groups[gi]
        dev = torch.device("cuda", self.device)
        had = gi in opt._hyper
        hyper = self.fused_hyper(opt, gi)
        if had and float(group["lr"]) != hyper["lr"]:
            # lr schedulers: refresh the device copy on the producer stream, ordered before the kernel
            hyper["dev"][1].fill_(float(group["lr"]))
            hyper["lr"] = float(group["lr"])
        if st is None:
            st = self.fused_state(opt, b)
        self._stream_wait_current(s)
        rc = lib.sdlc_fused_step(
            self._h, 1, b.sym.id, b.psym.id, b.numel, code, ctypes.c_void_p(st["s1"].data_ptr()),
            ctypes.c_void_p(st["s2"].data_ptr() if st["s2"] is not None else 0),
            ctypes.c_void_p(st["master"].data_ptr() if st["master"] is not None else 0),
            float(pre * post), ctypes.c_void_p(hyper["dev"].data_ptr()), s)
        if rc:
            raise CommError("fused optimizer step failed: %s" % self._err())
        hyper["advanced"] = False
        return lib.sdlc_record_event(self._h, s)

    def fused_state(self, opt, b):
        """This rank's shard of the optimizer state of bucket ``b`` (created on first use): fp32 momentum /
        Adam moments and, for 16-bit parameters, the fp32 master weights of the slice this rank owns."""
        if b.state is not None:
            return b.state
        lib = self._lib
        code = _TORCH_DTYPE[b.dtype]
        dev = torch.device("cuda", self.device)
        n = lib.sdlc_fused_state_elems(self._h, b.numel, code)
        st = {"s1": torch.zeros(n, dtype=torch.float32, device=dev), "s2": None, "master": None,
              "lo": lib.sdlc_fused_slice_start(self._h, b.numel, code), "n": n}
        if not isinstance(opt, torch.optim.SGD):
            st["s2"] = torch.zeros(n, dtype=torch.float32, device=dev)
        if b.dtype != torch.float32:
            m = torch.zeros(n, dtype=torch.float32, device=dev)
            avail = max(0, min(n, b.numel - st["lo"]))
            m[:avail] = b.pflat[st["lo"]:st["lo"] + avail].float()
            st["master"] = m
        b.state = st
        torch.cuda.current_stream(self.device).synchronize()
        return st

    def fused_hyper(self, opt, gi):
        hyper = opt._hyper.get(gi)
        if hyper is None:
            group = opt.param_groups[gi]
            vals = [float(v) for v in self.hyper_values(opt, group)] + [0.0] * 7
            hyper = {"dev": torch.tensor(vals, dtype=torch.float32, device=torch.device("cuda", self.device)),
                     "lr": float(group["lr"]), "advanced": False}
            opt._hyper[gi] = hyper
        return hyper

    @nvtx.annotate("sparkdl.fused_step_advance")
    def fused_step_advance(self, opt):
        """After the last bucket of an optimizer step: bump every group's device-side step counter (on the
        channel stream, behind the bucket kernels) and return an event the caller's stream must wait on."""
        lib = self._lib
        s = self.channel_stream(1)
        did = False
        for hyper in opt._hyper.values():
            if not hyper["advanced"]:
                if lib.sdlc_step_advance(self._h, 1, ctypes.c_void_p(hyper["dev"].data_ptr()), s):
                    raise CommError("step_advance failed")
                hyper["advanced"] = True
                did = True
        return lib.sdlc_record_event(self._h, s) if did else 0

    # ---------------------------------------------------------------------------------------------
    def close(self):
        if self._h:
            try:
                torch.cuda.synchronize(self.device)
            except Exception:  # pylint: disable=broad-except
                pass
            if self.world > 1 and self.ctx is not None:
                try:
                    store_barrier(self.ctx.store, "comm-close", self.rank, self.world, timeout=30.0)
                except Exception:  # pylint: disable=broad-except
                    pass
            self._lib.sdlc_comm_destroy(self._h)
            self._h = None
        if self._fd_server is not None:
            self._fd_server.close()
            self._fd_server = None


# -------------------------------------------------------------------------------------------------
# Single-process "virtual ranks" (tests / kernel development on ONE GPU): N communicators in one
# process whose peer pointers are ordinary device allocations of each other.
# -------------------------------------------------------------------------------------------------
class _FakeCtx(object):
    def __init__(self, rank, size):
        self.rank, self.size = rank, size
        self.store = None
        self.job_dir, self.job_id = "/tmp", "virt"


class VirtualGroup(object):
    """``world`` communicators on one device.  Kernels of different virtual ranks are launched on
    different streams and synchronise with each other exactly like real peers do."""

    def __init__(self, world, cfg=None, device=0, max_blocks=4):
        from spark_deep_learning_b200.utils.config import get_config
        cfg = cfg or get_config()
        self.world = world
        self.device = device
        torch.cuda.set_device(device)
        self.comms = [SymmComm(_FakeCtx(r, world), cfg, device=device, exchange=False) for r in range(world)]
        lib = self.comms[0]._lib
        for c in self.comms:
            lib.sdlc_comm_set_max_blocks(c._h, max_blocks)   # all ranks' CTAs must be co-resident
        self.pads = self.alloc(lib.sdlc_pad_bytes())
        for c, p in zip(self.comms, self.pads):
            lib.sdlc_comm_set_pads(c._h, p.id)
        if world > 1:
            self.ll = self.alloc(lib.sdlc_ll_bytes())
            for c, b in zip(self.comms, self.ll):
                lib.sdlc_comm_set_ll(c._h, b.id)
        self.streams = [torch.cuda.Stream(device=device) for _ in range(world)]

    def alloc(self, nbytes):
        lib = self.comms[0]._lib
        nbytes = int((nbytes + 255) // 256 * 256)
        bufs = []
        for c in self.comms:
            bid = lib.sdlc_buf_alloc(c._h, nbytes, 0)
            if bid < 0:
                raise CommError(c._err())
            bufs.append(SymmetricBuffer(c, bid, nbytes))
        for c, b in zip(self.comms, bufs):
            for r, other in enumerate(bufs):
                lib.sdlc_buf_set_peer_ptr(c._h, b.id, r, ctypes.c_void_p(other.ptr()))
        return bufs

    def close(self):
        torch.cuda.synchronize(self.device)
        for c in self.comms:
            c.ctx = None
            c.close()
