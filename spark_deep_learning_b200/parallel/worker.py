"""Worker-process entry point: ``python -m spark_deep_learning_b200.parallel.worker``.

One of these runs per task slot.  It unpickles ``(main, kwargs)`` with
cloudpickle (``/root/reference/sparkdl/horovod/runner_base.py:82-83``), calls ``main(**kwargs)``
and - on rank 0 - ships the pickled return value back to the driver
(``runner_base.py:93-95``).  A failure is written next to the payload as text so the driver
can re-raise it with the worker's traceback.
"""
import os
import sys
import threading
import time
import traceback


def _watch_driver(driver_pid):
    """Exit if the driver disappears (no orphaned GPU processes)."""
    while True:
        time.sleep(1.0)
        try:
            os.kill(driver_pid, 0)
        except ProcessLookupError:
            os._exit(113)
        except PermissionError:
            pass


def _maybe_inject_fault(rank, where):
    """``SPARKDL_FAULT=kill:rank=R:at=start|end`` or ``raise:rank=R``."""
    spec = os.environ.get("SPARKDL_FAULT", "")
    if not spec:
        return
    parts = spec.split(":")
    kind = parts[0]
    opts = dict(p.split("=", 1) for p in parts[1:] if "=" in p)
    if int(opts.get("rank", -1)) != rank or opts.get("at", "start") != where:
        return
    if kind == "kill":
        os._exit(int(opts.get("code", 137)))
    if kind == "raise":
        raise RuntimeError("injected fault on rank %d" % rank)
    if kind == "hang":
        time.sleep(float(opts.get("seconds", 3600)))


def _pin_device():
    """Bind this task slot to its GPU *before* user code runs (``runner_base.py:44-45``: a task slot
    "maps to a GPU"), so a ``main`` that touches CUDA before ``hvd.init()`` still lands on its own device.
    All GPUs stay visible (peer access needs them), so this is ``set_device``, not CUDA_VISIBLE_DEVICES."""
    if os.environ.get("SPARKDL_FORCE_CPU") == "1" or os.environ.get("SPARKDL_NO_PIN") == "1":
        return
    try:
        import torch
        if torch.cuda.is_available():
            idx = int(os.environ.get("SPARKDL_LOCAL_RANK", "0")) % torch.cuda.device_count()
            torch.cuda.set_device(idx)
            from spark_deep_learning_b200.utils import affinity
            affinity.bind_to_gpu(idx)          # threads + pinned host buffers on the GPU's own socket
    except Exception:  # pylint: disable=broad-except
        pass


def _arm_exit_watchdog(code, seconds=20.0):
    """Leave through the normal interpreter shutdown (at-exit hooks run, unlike ``os._exit``); a native
    detached thread ends the process if that shutdown wedges."""
    try:
        from spark_deep_learning_b200.parallel import core as _core
        _core.exit_watchdog(seconds, code)
    except Exception:  # pylint: disable=broad-except
        pass


def main():
    rank = int(os.environ["SPARKDL_RANK"])
    job_dir = os.environ["SPARKDL_JOB_DIR"]
    driver_pid = int(os.environ.get("SPARKDL_DRIVER_PID", "0"))
    if driver_pid:
        threading.Thread(target=_watch_driver, args=(driver_pid,), daemon=True).start()
    try:
        import cloudpickle
        _maybe_inject_fault(rank, "start")
        _pin_device()
        with open(os.path.join(job_dir, "payload.pkl"), "rb") as f:
            fn, kwargs = cloudpickle.load(f)
        result = fn(**kwargs)
        _maybe_inject_fault(rank, "end")
        if rank == 0:
            tmp = os.path.join(job_dir, "result.pkl.tmp")
            with open(tmp, "wb") as f:
                cloudpickle.dump(result, f)
            os.replace(tmp, os.path.join(job_dir, "result.pkl"))
        # orderly teardown of the collective runtime if the user forgot hvd.shutdown()
        try:
            hvd = sys.modules.get("spark_deep_learning_b200.parallel.hvd")
            if hvd is not None and hvd.is_initialized():
                hvd.shutdown()
        except Exception:  # pylint: disable=broad-except
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        return 0
    except BaseException:  # pylint: disable=broad-except
        tb = traceback.format_exc()
        try:
            with open(os.path.join(job_dir, "error.%d.txt" % rank), "w") as f:
                f.write(tb)
        except OSError:
            pass
        sys.stderr.write(tb)
        sys.stderr.flush()
        return 1


if __name__ == "__main__":
    code = main()
    sys.stdout.flush()
    sys.stderr.flush()
    _arm_exit_watchdog(code)
    sys.exit(code)
