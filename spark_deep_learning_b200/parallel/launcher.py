"""Local gang launcher behind :class:`sparkdl.HorovodRunner`.

Implements the behaviour the reference only documents:

* ``(main, kwargs)`` are serialized with cloudpickle and handed to every worker
  (``/root/reference/sparkdl/horovod/runner_base.py:82-83``);
* ``np < 0`` spawns ``-np`` subprocesses on this host (``runner_base.py:48``), one per task
  slot, slot *i* = GPU *i* on a GPU box or a CPU core otherwise (``runner_base.py:44-45``);
* worker stdout/stderr are multiplexed back to the driver and tee'd to a driver log file
  (``runner_base.py:49-51``), honouring ``driver_log_verbosity`` (``runner_base.py:62-72``);
* the value returned by rank 0 is the value of ``run`` (``runner_base.py:93-95``);
* gang semantics ("tasks starting all together", ``runner_base.py:54``): the first worker
  that dies takes the whole job down and its traceback is re-raised on the driver.
"""
import logging
import os
import signal
import subprocess
import sys
import tempfile
import threading
import time
import uuid

import cloudpickle

from spark_deep_learning_b200.parallel.store import StoreServer
from spark_deep_learning_b200.utils.config import get_config

MAX_LOG_MESSAGE_CHARS = 4000  # /root/reference/sparkdl/horovod/__init__.py:23

_logger = logging.getLogger("HorovodRunner")


class WorkerFailure(RuntimeError):
    """Raised on the driver when any rank exits abnormally."""

    def __init__(self, rank, returncode, traceback_text):
        self.rank = rank
        self.returncode = returncode
        self.traceback_text = traceback_text
        RuntimeError.__init__(
            self, "HorovodRunner worker rank %s failed (exit code %s)\n%s"
            % (rank, returncode, traceback_text or "<no traceback captured>"))


def available_slots():
    """Task slots on this box: GPUs if any are visible, otherwise CPU cores."""
    n = _gpu_count()
    if n > 0:
        return n, "gpu"
    try:
        return len(os.sched_getaffinity(0)), "cpu"
    except AttributeError:
        return os.cpu_count() or 1, "cpu"


def _gpu_count():
    override = os.environ.get("SPARKDL_NUM_GPUS")
    if override is not None:
        return int(override)
    # Do not create a CUDA context in the driver: count device nodes instead.
    try:
        n = len([d for d in os.listdir("/dev") if d.startswith("nvidia") and d[6:].isdigit()])
    except OSError:
        n = 0
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis is not None and n > 0:
        n = min(n, len([v for v in vis.split(",") if v.strip() != ""]))
    return n


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _LogMux(object):
    """Line-framed multiplexer of worker stdio with ``[rank]`` prefixes."""

    def __init__(self, log_path, stream_all):
        self._lock = threading.Lock()
        self._file = open(log_path, "a", buffering=1)
        self._stream_all = stream_all
        self.tail = {}

    def worker_line(self, rank, is_err, line):
        text = "[%d]%s %s" % (rank, "<stderr>" if is_err else "<stdout>", line)
        with self._lock:
            self._file.write(text + "\n")
            t = self.tail.setdefault(rank, [])
            t.append(line)
            if len(t) > 200:
                del t[:100]
            if self._stream_all:
                out = sys.stderr if is_err else sys.stdout
                out.write("[%d] %s\n" % (rank, line))
                out.flush()

    def driver_message(self, rank, message):
        """``log_to_driver`` payload: always printed on driver stdout."""
        message = str(message)[:MAX_LOG_MESSAGE_CHARS]
        with self._lock:
            self._file.write("[%s]<log_to_driver> %s\n" % (rank, message))
            sys.stdout.write(message + "\n")
            sys.stdout.flush()

    def close(self):
        with self._lock:
            self._file.close()


def _pump(pipe, rank, is_err, mux):
    try:
        for raw in iter(pipe.readline, b""):
            mux.worker_line(rank, is_err, raw.decode("utf-8", "replace").rstrip("\n"))
    except (OSError, ValueError):
        pass
    finally:
        try:
            pipe.close()
        except OSError:
            pass


def _kill_group(proc, sig):
    try:
        os.killpg(proc.pid, sig)
    except (ProcessLookupError, PermissionError, OSError):
        try:
            proc.send_signal(sig)
        except (ProcessLookupError, OSError):
            pass


def launch_gang(main, kwargs, num_proc, stream_all_logs, extra_env=None):
    """Run ``main(**kwargs)`` on ``num_proc`` local worker processes; return rank 0's value."""
    cfg = get_config(refresh=True)
    job_id = uuid.uuid4().hex[:12]
    job_dir = tempfile.mkdtemp(prefix="sparkdl-job-%s-" % job_id)
    payload_path = os.path.join(job_dir, "payload.pkl")
    with open(payload_path, "wb") as f:
        cloudpickle.dump((main, kwargs), f)

    log_path = os.environ.get("SPARKDL_DRIVER_LOG") or os.path.join(job_dir, "driver.log")
    mux = _LogMux(log_path, stream_all_logs)
    store = StoreServer(os.path.join(job_dir, "store.sock"), log_sink=mux.driver_message)
    _, slot_kind = available_slots()
    master_port = _free_port()

    procs, pumps = [], []
    try:
        for rank in range(num_proc):
            env = dict(os.environ)
            env.update({
                "SPARKDL_RANK": str(rank), "SPARKDL_SIZE": str(num_proc),
                "SPARKDL_LOCAL_RANK": str(rank), "SPARKDL_LOCAL_SIZE": str(num_proc),
                "SPARKDL_JOB_ID": job_id, "SPARKDL_JOB_DIR": job_dir,
                "SPARKDL_STORE": store.path, "SPARKDL_SLOT_KIND": slot_kind,
                "SPARKDL_DRIVER_PID": str(os.getpid()),
                # Horovod / torch.distributed spellings so unmodified mains find their peers
                "HOROVOD_RANK": str(rank), "HOROVOD_SIZE": str(num_proc),
                "HOROVOD_LOCAL_RANK": str(rank), "HOROVOD_LOCAL_SIZE": str(num_proc),
                "RANK": str(rank), "WORLD_SIZE": str(num_proc), "LOCAL_RANK": str(rank),
                "LOCAL_WORLD_SIZE": str(num_proc),
                "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(master_port),
                "PYTHONUNBUFFERED": "1",
            })
            env.pop("SPARKDL_EXTERNAL_GANG", None)
            # make this package importable in the worker even when it is not installed
            root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
            if os.environ.get("SPARKDL_HOROVOD_SHIM") == "1":
                env["PYTHONPATH"] = os.path.join(root, "spark_deep_learning_b200", "compat") + os.pathsep + env["PYTHONPATH"]
            if extra_env:
                env.update(extra_env)
            p = subprocess.Popen(
                [sys.executable, "-u", "-m", "spark_deep_learning_b200.parallel.worker"],
                env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                stdin=subprocess.DEVNULL, start_new_session=True, cwd=os.getcwd())
            procs.append(p)
            for pipe, is_err in ((p.stdout, False), (p.stderr, True)):
                t = threading.Thread(target=_pump, args=(pipe, rank, is_err, mux), daemon=True)
                t.start()
                pumps.append(t)

        failed = _wait_all(procs, store, cfg)
        for t in pumps:
            t.join(timeout=5.0)
        if failed is not None:
            rank, rc = failed
            tb = _read_error(job_dir, rank)
            if tb is None:
                tb = "\n".join(mux.tail.get(rank, [])[-40:])
            raise WorkerFailure(rank, rc, tb)
        result_path = os.path.join(job_dir, "result.pkl")
        if not os.path.exists(result_path):
            raise WorkerFailure(0, 0, "rank 0 exited without producing a result")
        with open(result_path, "rb") as f:
            return cloudpickle.load(f)
    finally:
        for p in procs:
            if p.poll() is None:
                _kill_group(p, signal.SIGKILL)
        for p in procs:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                pass
        store.close()
        mux.close()
        if not os.environ.get("SPARKDL_KEEP_JOB_DIR"):
            _cleanup_job(job_dir, job_id)


def _wait_all(procs, store, cfg):
    """Block until every worker exits.  Returns ``None`` or ``(rank, returncode)`` of the
    first failure; on failure the surviving siblings are terminated (gang semantics)."""
    pending = set(range(len(procs)))
    failed = None
    while pending:
        for r in list(pending):
            rc = procs[r].poll()
            if rc is None:
                continue
            pending.discard(r)
            if rc != 0 and failed is None:
                failed = (r, rc)
        if failed is not None and pending:
            store.set("abort", True)
            deadline = time.time() + 3.0
            for r in pending:
                _kill_group(procs[r], signal.SIGTERM)
            while time.time() < deadline and any(procs[r].poll() is None for r in pending):
                time.sleep(0.05)
            for r in pending:
                if procs[r].poll() is None:
                    _kill_group(procs[r], signal.SIGKILL)
            for r in pending:
                try:
                    procs[r].wait(timeout=10)
                except subprocess.TimeoutExpired:
                    pass
            pending.clear()
        if pending:
            time.sleep(0.01)
    return failed


def _read_error(job_dir, rank):
    path = os.path.join(job_dir, "error.%d.txt" % rank)
    if os.path.exists(path):
        with open(path, "r") as f:
            return f.read()
    return None


def _cleanup_job(job_dir, job_id):
    import shutil
    shutil.rmtree(job_dir, ignore_errors=True)
    # shared-memory segments created by the CPU data plane for this job
    try:
        for name in os.listdir("/dev/shm"):
            if name.startswith("sparkdl-%s" % job_id):
                try:
                    os.unlink(os.path.join("/dev/shm", name))
                except OSError:
                    pass
    except OSError:
        pass


def resolve_num_proc(np):
    """Map the documented ``np`` values onto a local process count.

    * ``np < 0``  -> ``-np`` local subprocesses (``runner_base.py:48``).
    * ``np > 0``  -> cluster mode.  There is no Spark cluster on this box, so the local task
      slots play the cluster: the job "will fail" if ``np`` exceeds them (``runner_base.py:57-58``).
    * ``np == 0`` -> deprecated "use all task slots" (``/root/reference/README.md:57-61``).
    """
    if not isinstance(np, int) or isinstance(np, bool):
        raise TypeError("np must be an int, got %r" % (np,))
    slots, kind = available_slots()
    if np < 0:
        n = -np
        if kind == "gpu" and n > slots:
            raise RuntimeError(
                "HorovodRunner(np=%d) needs %d GPU task slots but this host has %d"
                % (np, n, slots))
        return n
    if np == 0:
        _logger.warning("HorovodRunner(np=0) is deprecated; using all %d %s task slots.",
                        slots, kind)
        return slots
    if np > slots:
        raise RuntimeError(
            "HorovodRunner(np=%d): only %d %s task slots are available on this cluster; "
            "the job fails as documented." % (np, slots, kind))
    return np
