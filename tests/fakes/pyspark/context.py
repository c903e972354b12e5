import os
import subprocess
import sys
import tempfile
import threading
from multiprocessing.connection import Client, Listener

import cloudpickle


class BarrierTaskInfo(object):
    def __init__(self, address):
        self.address = address


class BarrierTaskContext(object):
    _current = None

    def __init__(self, pid, n, addr):
        self._pid, self._n, self._addr, self._epoch = pid, n, addr, 0

    @classmethod
    def get(cls):
        if cls._current is None:
            raise RuntimeError("BarrierTaskContext.get() outside of a barrier task")
        return cls._current

    def partitionId(self):
        return self._pid

    def getTaskInfos(self):
        return [BarrierTaskInfo("localhost:%d" % (40000 + i)) for i in range(self._n)]

    def allGather(self, message=""):
        self._epoch += 1
        with Client(self._addr, family="AF_UNIX") as c:
            c.send((self._epoch, self._pid, message))
            return c.recv()

    def barrier(self):
        self.allGather("")


class _Coordinator(threading.Thread):
    """allGather / barrier rendezvous of one barrier stage."""

    def __init__(self, n):
        super().__init__(daemon=True)
        self.n = n
        self.dir = tempfile.mkdtemp(prefix="fakespark-")
        self.addr = os.path.join(self.dir, "coord.sock")
        self.listener = Listener(self.addr, family="AF_UNIX")
        self.rounds = {}
        self.stop = False

    def run(self):
        while not self.stop:
            try:
                conn = self.listener.accept()
            except OSError:
                return
            epoch, pid, msg = conn.recv()
            waiting = self.rounds.setdefault(epoch, {})
            waiting[pid] = (conn, msg)
            if len(waiting) == self.n:
                msgs = [waiting[i][1] for i in range(self.n)]
                for c, _ in waiting.values():
                    c.send(msgs)
                    c.close()
                del self.rounds[epoch]

    def close(self):
        self.stop = True
        try:
            self.listener.close()
        except OSError:
            pass


_WORKER = r'''
import sys, cloudpickle
sys.path[:0] = %r
from pyspark.context import BarrierTaskContext
pid, n, addr, fpath, opath = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
if addr != "-":
    BarrierTaskContext._current = BarrierTaskContext(pid, n, addr)
with open(fpath, "rb") as f:
    func, rows = cloudpickle.load(f)
out = list(func(iter(rows)))
with open(opath, "wb") as f:
    cloudpickle.dump(out, f)
'''


class RDD(object):
    def __init__(self, sc, partitions, func=None, barrier=False):
        self._sc, self._parts, self._func, self._barrier = sc, partitions, func, barrier

    def getNumPartitions(self):
        return len(self._parts)

    def barrier(self):
        return RDDBarrier(self)

    def mapPartitions(self, f, preservesPartitioning=False):  # pylint: disable=unused-argument
        prev = self._func
        func = f if prev is None else (lambda it, prev=prev, f=f: f(prev(it)))
        return RDD(self._sc, self._parts, func, self._barrier)

    def collect(self):
        if self._func is None:
            return [r for p in self._parts for r in p]
        n = len(self._parts)
        if self._barrier and n > self._sc.defaultParallelism:
            raise RuntimeError("[SPARK-24819]: Barrier execution mode does not allow run a barrier stage that requires "
                               "more slots than the total number of slots in the cluster currently (%d > %d)."
                               % (n, self._sc.defaultParallelism))
        coord = _Coordinator(n) if self._barrier else None
        if coord:
            coord.start()
        work = tempfile.mkdtemp(prefix="fakespark-stage-")
        procs = []
        try:
            for i, rows in enumerate(self._parts):
                fpath, opath = os.path.join(work, "f%d.pkl" % i), os.path.join(work, "o%d.pkl" % i)
                with open(fpath, "wb") as f:
                    cloudpickle.dump((self._func, list(rows)), f)
                cmd = [sys.executable, "-c", _WORKER % (list(sys.path),), str(i), str(n), coord.addr if coord else "-", fpath, opath]
                procs.append((subprocess.Popen(cmd, stderr=subprocess.PIPE, text=True), opath))
            out = []
            errs = []
            for p, opath in procs:
                _, err = p.communicate(timeout=600)
                if p.returncode != 0:
                    errs.append(err[-2000:])
                else:
                    sys.stderr.write(err)
                    with open(opath, "rb") as f:
                        out.extend(cloudpickle.load(f))
            if errs:
                for p, _ in procs:
                    if p.poll() is None:
                        p.kill()
                raise RuntimeError("barrier stage failed:\n" + "\n".join(errs))
            return out
        finally:
            if coord:
                coord.close()


class RDDBarrier(object):
    def __init__(self, rdd):
        self._rdd = rdd

    def mapPartitions(self, f, preservesPartitioning=False):  # pylint: disable=unused-argument
        r = self._rdd.mapPartitions(f)
        r._barrier = True
        return r


class SparkContext(object):
    def __init__(self, slots):
        self.defaultParallelism = slots

    def parallelize(self, c, numSlices=None):
        items = list(c)
        n = numSlices or self.defaultParallelism
        k, m = divmod(len(items), n)
        parts, s = [], 0
        for i in range(n):
            e = s + k + (1 if i < m else 0)
            parts.append(items[s:e])
            s = e
        return RDD(self, parts)
