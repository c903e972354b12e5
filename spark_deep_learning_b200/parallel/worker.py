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


def main():
    rank = int(os.environ["SPARKDL_RANK"])
    job_dir = os.environ["SPARKDL_JOB_DIR"]
    driver_pid = int(os.environ.get("SPARKDL_DRIVER_PID", "0"))
    if driver_pid:
        threading.Thread(target=_watch_driver, args=(driver_pid,), daemon=True).start()
    try:
        import cloudpickle
        _maybe_inject_fault(rank, "start")
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
    os._exit(code)
