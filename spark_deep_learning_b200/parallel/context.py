"""Per-process job context: who am I, and how do I reach the driver / my peers.

Resolves the three bootstrap modes (SURVEY.md §5.8 "Bootstrap"):

1. spawned by :mod:`launcher`   - ``SPARKDL_RANK/SIZE/STORE`` are set;
2. external gang (torchrun)     - ``RANK/WORLD_SIZE/MASTER_PORT`` are set: rank 0 hosts the
   store on a unix socket whose name is derived from ``MASTER_PORT``;
3. plain single process         - size 1, in-memory store.
"""
import atexit
import os
import stat
import threading

from spark_deep_learning_b200.parallel.store import (LocalStore, NamespacedStore, StoreClient, StoreServer)

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
            # The store speaks pickle, so it must only be reachable by this user: a 0700 directory owned by us
            # (under $XDG_RUNTIME_DIR when there is one) holds the socket and the fd-passing sockets.
            base = env.get("XDG_RUNTIME_DIR") or "/tmp"
            self.job_dir = os.path.join(base, "sparkdl-ext-%d-%s" % (os.getuid(), tag))
            path = os.path.join(self.job_dir, "store.sock")
            if self.rank == 0:
                _private_dir(self.job_dir)
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
        if getattr(self, "mode", None) == "external" and self.rank == 0 and \
                os.path.basename(self.job_dir).startswith("sparkdl-ext-"):
            import shutil
            shutil.rmtree(self.job_dir, ignore_errors=True)     # created by this rank (_private_dir)
        if self._server is not None:
            self._server.close()
            self._server = None


def _private_dir(path):
    """Create ``path`` with mode 0700, or insist that an existing one is ours and not group/world accessible."""
    try:
        os.mkdir(path, 0o700)
    except FileExistsError:
        st = os.lstat(path)
        if not stat.S_ISDIR(st.st_mode) or st.st_uid != os.getuid():
            raise RuntimeError("refusing to use %s: not a directory owned by uid %d" % (path, os.getuid()))
        if st.st_mode & 0o077:
            os.chmod(path, 0o700)


def get_context():
    """The process-wide job context.  It outlives ``hvd.shutdown()`` (the store connection is the only way
    back to the driver / rank 0, and a later ``hvd.init()`` needs it); ``atexit`` closes it."""
    global _ctx
    with _lock:
        if _ctx is None:
            _ctx = JobContext()
            atexit.register(reset_context)
        return _ctx


def generation_store(ctx, generation):
    """The job store seen through the key namespace of this ``hvd.init()`` generation."""
    return NamespacedStore(ctx.store, "g%d/" % generation)


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
