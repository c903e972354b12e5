"""Typed runtime configuration, read once per process from the environment.

The reference has no flag system beyond constructor kwargs (SURVEY.md §5.6;
``/root/reference/sparkdl/horovod/runner_base.py:39``).  Stock Horovod is driven
by ``HOROVOD_*`` environment variables, so every knob here accepts both the
``SPARKDL_*`` spelling and the common ``HOROVOD_*`` spelling.  The launcher
copies the driver's environment into every worker, so all ranks agree.
"""
import os


def _env(names, default, cast):
    for n in names:
        v = os.environ.get(n)
        if v is not None and v != "":
            try:
                return cast(v)
            except (TypeError, ValueError):
                raise ValueError("bad value %r for environment variable %s" % (v, n))
    return default


def _bool(v):
    return str(v).strip().lower() in ("1", "true", "yes", "on")


class Config(object):
    """Snapshot of all tunables.  Construct with :func:`get_config`."""

    def __init__(self):
        # --- engine / fusion (Horovod names honoured) -------------------------------------
        self.fusion_threshold = _env(
            ["SPARKDL_FUSION_THRESHOLD", "HOROVOD_FUSION_THRESHOLD"], 64 << 20, int)
        self.cycle_time_ms = _env(["SPARKDL_CYCLE_TIME", "HOROVOD_CYCLE_TIME"], 0.5, float)
        self.timeline = _env(["SPARKDL_TIMELINE", "HOROVOD_TIMELINE"], "", str)
        self.stall_check_s = _env(
            ["SPARKDL_STALL_CHECK_TIME_SECONDS", "HOROVOD_STALL_CHECK_TIME_SECONDS"], 60.0, float)
        self.stall_shutdown_s = _env(
            ["SPARKDL_STALL_SHUTDOWN_TIME_SECONDS", "HOROVOD_STALL_SHUTDOWN_TIME_SECONDS"],
            0.0, float)
        self.autotune = _env(["SPARKDL_AUTOTUNE", "HOROVOD_AUTOTUNE"], False, _bool)
        # response cache: repeated (name, shape, dtype, op) collectives are announced with one bit (0 disables, like
        # HOROVOD_CACHE_CAPACITY=0)
        self.response_cache = _env(["SPARKDL_RESPONSE_CACHE", "HOROVOD_CACHE_CAPACITY"], 1024, int) > 0
        # --- CPU shared-memory data plane --------------------------------------------------
        self.shm_stage_bytes = _env(["SPARKDL_SHM_STAGE_BYTES"], 8 << 20, int)
        # --- GPU collectives ---------------------------------------------------------------
        # message size (bytes) at or below which the one-shot kernel is used
        self.oneshot_max_bytes = _env(["SPARKDL_ONESHOT_MAX_BYTES"], 16 << 10, int)   # measured crossover, profiles/
        # use NVLS multicast (multimem.*) when the hardware offers it
        self.use_multicast = _env(["SPARKDL_USE_MULTICAST"], True, _bool)
        # symmetric-memory transport: auto | vmm | ipc
        self.transport = _env(["SPARKDL_TRANSPORT"], "auto", str)
        # CTAs used by a bandwidth-regime collective (kept small so backward keeps its SMs)
        self.comm_ctas = _env(["SPARKDL_COMM_CTAS"], 32, int)
        # bucket size for DistributedOptimizer (bytes of gradient per fused launch)
        self.bucket_bytes = _env(["SPARKDL_BUCKET_BYTES"], 32 << 20, int)
        # in-kernel spin-wait timeout before a collective aborts (milliseconds)
        self.kernel_timeout_ms = _env(["SPARKDL_KERNEL_TIMEOUT_MS"], 20000, int)
        # wire format for gradient buckets: native | bf16 | fp16 | fp8
        self.wire_dtype = _env(["SPARKDL_WIRE_DTYPE"], "native", str)
        # --- fault injection (tests) ---------------------------------------
        self.fault = _env(["SPARKDL_FAULT"], "", str)
        # --- launcher ------------------------------------------------------------------------
        self.start_timeout_s = _env(
            ["SPARKDL_START_TIMEOUT", "HOROVOD_START_TIMEOUT"], 120.0, float)
        self.inprocess_single = _env(["SPARKDL_INPROCESS_SINGLE"], False, _bool)

    def as_dict(self):
        return dict(self.__dict__)


_CONFIG = None


def get_config(refresh=False):
    global _CONFIG
    if _CONFIG is None or refresh:
        _CONFIG = Config()
    return _CONFIG
