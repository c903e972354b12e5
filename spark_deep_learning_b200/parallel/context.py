"""Per-process job context: who am I, and how do I reach the driver / my peers.

Resolves the three bootstrap modes (SURVEY.md §5.8 "Bootstrap"):

1. spawned by :mod:`launcher`   - ``SPARKDL_RANK/SIZE/STORE`` are set;
2. external gang (torchrun)     - ``RANK/WORLD_SIZE/MASTER_PORT`` are set: rank 0 hosts the
   store on a unix socket whose name is derived from ``MASTER_PORT``;
3. plain single process         - size 1, in-memory store.
"""
import os
import threading

from spark_deep_learning_b200.parallel.store import (LocalStore, StoreClient, StoreServer)

_lock = threading.Lock()
_ctx = None


class JobContext(object):
    def __init__(self):
        env = os.environ
        self._server = None
        if "SPARKDL_RANK" in env and "SPARKDL_STORE" in env:
            self.mode = "launcher"
            self.rank = int(env["SPARKDL_RANK"])
            self.size = int(env["SPARKDL_SIZE"])
            self.local_rank = int(env.get("SPARKDL_LOCAL_RANK", self.rank))
            self.local_size = int(env.get("SPARKDL_LOCAL_SIZE", self.size))
            self.job_id = env.get("SPARKDL_JOB_ID", "job")
            self.job_dir = env.get("SPARKDL_JOB_DIR", "/tmp")
            self.store = StoreClient(env["SPARKDL_STORE"])
        elif "RANK" in env and "WORLD_SIZE" in env and int(env["WORLD_SIZE"]) > 1:
            self.mode = "external"
            self.rank = int(env["RANK"])
            self.size = int(env["WORLD_SIZE"])
            self.local_rank = int(env.get("LOCAL_RANK", self.rank))
            self.local_size = int(env.get("LOCAL_WORLD_SIZE", self.size))
            tag = "%s-%s" % (env.get("MASTER_PORT", "0"),
                             env.get("TORCHELASTIC_RUN_ID", "none"))
            tag = "".join(c if c.isalnum() or c in "-_" else "_" for c in tag)[:48]
            self.job_id = "ext" + tag
            self.job_dir = "/tmp/sparkdl-ext-%s" % tag
            path = "/tmp/sparkdl-store-%s.sock" % tag
            if self.rank == 0:
                os.makedirs(self.job_dir, exist_ok=True)
                self._server = StoreServer(path)
            self.store = StoreClient(path)
            # Every rank must see the job dir before using it for fd sockets.
            if self.rank == 0:
                self.store.set("jobdir_ready", True)
            else:
                self.store.get("jobdir_ready")
        else:
            self.mode = "single"
            self.rank = 0
            self.size = 1
            self.local_rank = 0
            self.local_size = 1
            self.job_id = "single%d" % os.getpid()
            self.job_dir = "/tmp"
            self.store = LocalStore()

    def close(self):
        try:
            self.store.close()
        except Exception:  # pylint: disable=broad-except
            pass
        if self._server is not None:
            self._server.close()
            self._server = None


def get_context():
    global _ctx
    with _lock:
        if _ctx is None:
            _ctx = JobContext()
        return _ctx


def reset_context():
    global _ctx
    with _lock:
        if _ctx is not None:
            _ctx.close()
        _ctx = None


def send_log_to_driver(message):
    """Deliver one ``log_to_driver`` message (already truncated) to the driver's stdout."""
    env = os.environ
    if "SPARKDL_STORE" in env and "SPARKDL_RANK" in env:
        ctx = get_context()
        ctx.store.log(ctx.rank, message)
    else:
        # in-process / external gang: the "driver" is this process's stdout
        print(message, flush=True)
