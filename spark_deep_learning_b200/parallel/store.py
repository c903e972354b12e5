"""Bootstrap key-value store over a unix-domain socket.

This is the rendezvous layer that the reference only implies ("launches the Horovod job",
``/root/reference/sparkdl/horovod/runner_base.py:27``; SURVEY.md §2e "Rendezvous / bootstrap:
absent").  Everything is on one box, so a unix socket is enough and needs no network:

* launcher mode   - the driver hosts the store at ``<jobdir>/store.sock``;
* external gangs  - rank 0 hosts it at a path derived from ``MASTER_PORT`` (torchrun);
* single process  - an in-memory dict, no socket at all.

Besides ``set/get/add/wait`` the same socket carries the worker -> driver control messages
of the logging contract (``log_to_driver``, ``/root/reference/sparkdl/horovod/__init__.py:20-25``)
and lets ranks hand CUDA VMM file descriptors to each other with ``SCM_RIGHTS``.
"""
import os
import pickle
import socket
import struct
import threading
import time

_HDR = struct.Struct("!I")


def _send_msg(sock, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(_HDR.pack(len(data)) + data)


def _recv_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("store peer closed the connection")
        buf.extend(chunk)
    return bytes(buf)


def _recv_msg(sock):
    (n,) = _HDR.unpack(_recv_exact(sock, _HDR.size))
    return pickle.loads(_recv_exact(sock, n))


class StoreServer(object):
    """Threaded KV server.  ``log_sink(rank, message)`` receives ``log_to_driver`` messages."""

    def __init__(self, path, log_sink=None):
        self.path = path
        self._kv = {}
        self._cv = threading.Condition()
        self._log_sink = log_sink
        self._closed = False
        try:
            os.unlink(path)
        except OSError:
            pass
        self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._sock.bind(path)
        self._sock.listen(128)
        self._thread = threading.Thread(target=self._accept_loop, name="sparkdl-store", daemon=True)
        self._thread.start()

    # -- direct (in-process) access, used by the driver itself ------------------------------
    def set(self, key, value):
        with self._cv:
            self._kv[key] = value
            self._cv.notify_all()

    def get(self, key, timeout=None):
        deadline = None if timeout is None else time.time() + timeout
        with self._cv:
            while key not in self._kv:
                left = None if deadline is None else deadline - time.time()
                if left is not None and left <= 0:
                    raise TimeoutError("store key %r not set within %.1fs" % (key, timeout))
                self._cv.wait(left if left is not None else 1.0)
                if self._closed:
                    raise ConnectionError("store closed")
            return self._kv[key]

    def try_get(self, key, default=None):
        with self._cv:
            return self._kv.get(key, default)

    def add(self, key, amount):
        with self._cv:
            v = self._kv.get(key, 0) + amount
            self._kv[key] = v
            self._cv.notify_all()
            return v

    # -- socket side ---------------------------------------------------------------------------
    def _accept_loop(self):
        while not self._closed:
            try:
                conn, _ = self._sock.accept()
            except OSError:
                return
            t = threading.Thread(target=self._serve, args=(conn,), daemon=True)
            t.start()

    def _serve(self, conn):
        try:
            while True:
                req = _recv_msg(conn)
                op = req[0]
                if op == "set":
                    self.set(req[1], req[2])
                    _send_msg(conn, ("ok", None))
                elif op == "get":
                    try:
                        _send_msg(conn, ("ok", self.get(req[1], req[2])))
                    except TimeoutError as e:
                        _send_msg(conn, ("timeout", str(e)))
                elif op == "try_get":
                    _send_msg(conn, ("ok", self.try_get(req[1])))
                elif op == "add":
                    _send_msg(conn, ("ok", self.add(req[1], req[2])))
                elif op == "log":
                    if self._log_sink is not None:
                        self._log_sink(req[1], req[2])
                    _send_msg(conn, ("ok", None))
                elif op == "close":
                    _send_msg(conn, ("ok", None))
                    return
                else:
                    _send_msg(conn, ("error", "unknown op %r" % (op,)))
        except (ConnectionError, OSError, EOFError):
            pass
        finally:
            try:
                conn.close()
            except OSError:
                pass

    def close(self):
        self._closed = True
        with self._cv:
            self._cv.notify_all()
        try:
            self._sock.close()
        except OSError:
            pass
        try:
            os.unlink(self.path)
        except OSError:
            pass


class StoreClient(object):
    """Blocking client.  One request in flight at a time (guarded by a lock)."""

    def __init__(self, path, connect_timeout=120.0):
        self.path = path
        self._lock = threading.Lock()
        deadline = time.time() + connect_timeout
        last = None
        while True:
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            try:
                s.connect(path)
                self._sock = s
                break
            except OSError as e:  # server not up yet (external-gang mode) - retry
                s.close()
                last = e
                if time.time() > deadline:
                    raise TimeoutError("cannot reach sparkdl store at %s: %s" % (path, last))
                time.sleep(0.02)

    def _call(self, *req):
        with self._lock:
            _send_msg(self._sock, req)
            status, val = _recv_msg(self._sock)
        if status == "ok":
            return val
        if status == "timeout":
            raise TimeoutError(val)
        raise RuntimeError(val)

    def set(self, key, value):
        self._call("set", key, value)

    def get(self, key, timeout=300.0):
        return self._call("get", key, timeout)

    def try_get(self, key, default=None):
        v = self._call("try_get", key)
        return default if v is None else v

    def add(self, key, amount=1):
        return self._call("add", key, amount)

    def log(self, rank, message):
        self._call("log", rank, message)

    def close(self):
        try:
            self._call("close")
        except Exception:  # pylint: disable=broad-except
            pass
        try:
            self._sock.close()
        except OSError:
            pass


class LocalStore(object):
    """Single-process stand-in with the same interface (size == 1, no launcher)."""

    def __init__(self, log_sink=None):
        self._kv = {}
        self._log_sink = log_sink

    def set(self, key, value):
        self._kv[key] = value

    def get(self, key, timeout=None):
        if key not in self._kv:
            raise TimeoutError("key %r not set (single-process store)" % (key,))
        return self._kv[key]

    def try_get(self, key, default=None):
        return self._kv.get(key, default)

    def add(self, key, amount=1):
        self._kv[key] = self._kv.get(key, 0) + amount
        return self._kv[key]

    def log(self, rank, message):
        if self._log_sink is not None:
            self._log_sink(rank, message)
        else:
            print(message, flush=True)

    def close(self):
        pass


class NamespacedStore(object):
    """View of a store with every key prefixed.  ``hvd.init()`` wraps the job store in
    ``NamespacedStore(store, "g<N>/")`` with N = how many times this process has initialised, so a
    ``shutdown()`` -> ``init()`` cycle never reads the bootstrap keys (shm segment name, IPC handles, votes,
    barrier counters) of the previous incarnation."""

    def __init__(self, base, prefix):
        self._base = base
        self._prefix = prefix

    def set(self, key, value):
        return self._base.set(self._prefix + key, value)

    def get(self, key, *args, **kwargs):
        return self._base.get(self._prefix + key, *args, **kwargs)

    def try_get(self, key, default=None):
        return self._base.try_get(self._prefix + key, default)

    def add(self, key, amount=1):
        return self._base.add(self._prefix + key, amount)

    def log(self, rank, message):
        return self._base.log(rank, message)

    def close(self):
        pass   # the base store belongs to the job context


def store_barrier(store, name, rank, size, timeout=300.0):
    """Counter barrier on the store (cold path only: init / shutdown / handle exchange)."""
    if size <= 1:
        return
    n = store.add("barrier/%s/count" % name, 1)
    if n == size:
        store.set("barrier/%s/done" % name, True)
    store.get("barrier/%s/done" % name, timeout)


# ---------------------------------------------------------------------------------------------
# File-descriptor passing (CUDA VMM POSIX handles) between ranks of one box.
# ---------------------------------------------------------------------------------------------
class FdServer(object):
    """Serves a mutable table ``{key: fd}`` to peers through ``SCM_RIGHTS``."""

    def __init__(self, path):
        self.path = path
        self._fds = {}
        self._cv = threading.Condition()
        self._closed = False
        try:
            os.unlink(path)
        except OSError:
            pass
        self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._sock.bind(path)
        self._sock.listen(64)
        self._thread = threading.Thread(target=self._loop, name="sparkdl-fdserver", daemon=True)
        self._thread.start()

    def publish(self, key, fd):
        with self._cv:
            self._fds[key] = fd
            self._cv.notify_all()

    def _loop(self):
        while not self._closed:
            try:
                conn, _ = self._sock.accept()
            except OSError:
                return
            threading.Thread(target=self._serve, args=(conn,), daemon=True).start()

    def _serve(self, conn):
        try:
            while True:
                key = _recv_msg(conn)
                if key is None:
                    return
                with self._cv:
                    deadline = time.time() + 120.0
                    while key not in self._fds and time.time() < deadline and not self._closed:
                        self._cv.wait(0.5)
                    fd = self._fds.get(key)  # None = the owner failed to export it
                if fd is None:
                    socket.send_fds(conn, [b"N"], [])
                else:
                    socket.send_fds(conn, [b"Y"], [fd])
        except (ConnectionError, OSError, EOFError):
            pass
        finally:
            conn.close()

    def close(self):
        self._closed = True
        with self._cv:
            self._cv.notify_all()
        try:
            self._sock.close()
        except OSError:
            pass
        try:
            os.unlink(self.path)
        except OSError:
            pass


def fetch_fd(path, key, timeout=120.0):
    """Connect to a peer's :class:`FdServer` and receive the descriptor published as ``key``."""
    deadline = time.time() + timeout
    while True:
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        try:
            s.connect(path)
            break
        except OSError:
            s.close()
            if time.time() > deadline:
                raise TimeoutError("cannot reach fd server %s" % path)
            time.sleep(0.02)
    try:
        _send_msg(s, key)
        msg, fds, _, _ = socket.recv_fds(s, 16, 4)
        _send_msg(s, None)
        if msg != b"Y" or not fds:
            raise RuntimeError("peer did not publish fd %r" % (key,))
        return fds[0]
    finally:
        s.close()
