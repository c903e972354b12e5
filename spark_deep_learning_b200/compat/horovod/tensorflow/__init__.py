"""``horovod.tensorflow``-style frontend over the sm_100a / shm collective runtime.

The reference's only framework-specific module is ``sparkdl.horovod.tensorflow.keras``
(``/root/reference/sparkdl/horovod/tensorflow/keras.py:9-11``), i.e. its users write ``horovod.tensorflow.keras``
mains.  TensorFlow is not installable on the B200 image, so this frontend is framework-neutral at its core: tensors
cross into the runtime as NumPy arrays (``.numpy()`` of an eager tf.Tensor / tf.Variable, or anything ``np.asarray``
accepts) and come back in the caller's type (``tf.convert_to_tensor`` when TensorFlow is importable).  The collective
itself is the same engine + kernels as the torch frontend (``spark_deep_learning_b200/parallel/hvd.py``).
"""
import numpy as np
import torch

from spark_deep_learning_b200.parallel import hvd as _hvd
from spark_deep_learning_b200.parallel.hvd import (Average, Max, Min, Product, Sum, barrier, init,  # noqa: F401
                                                   is_initialized, local_rank, local_size, rank, shutdown, size)

__all__ = ["init", "shutdown", "rank", "size", "local_rank", "local_size", "allreduce", "allgather", "broadcast",
           "broadcast_variables", "broadcast_object", "allgather_object", "grouped_allreduce", "DistributedGradientTape",
           "Average", "Sum", "Min", "Max", "Product", "barrier"]


def _tf():
    try:
        import tensorflow as tf  # pylint: disable=import-error
        return tf
    except Exception:  # pylint: disable=broad-except
        return None


def _to_numpy(x):
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        x = x.numpy()
    return np.ascontiguousarray(np.asarray(x))


def _like(out, ref):
    """Return ``out`` (ndarray) in the type the caller handed in."""
    tf = _tf()
    if tf is not None and isinstance(ref, (tf.Tensor, tf.Variable)):
        return tf.convert_to_tensor(out)
    if isinstance(ref, (float, int)) or (isinstance(ref, np.ndarray) and ref.ndim == 0) or np.isscalar(ref):
        return out.reshape(()).item() if not isinstance(ref, np.ndarray) else out.reshape(())
    return out


def _torch(a):
    t = torch.from_numpy(a.reshape(-1) if a.ndim == 0 else a)
    return t


def allreduce(tensor, average=None, name=None, op=None, prescale_factor=1.0, postscale_factor=1.0):
    """Reduce ``tensor`` over all workers (default: average), same result on every worker."""
    a = _to_numpy(tensor)
    out = _hvd.allreduce(_torch(a), average=average, name=name, op=op, prescale_factor=prescale_factor,
                         postscale_factor=postscale_factor)
    return _like(out.numpy().reshape(a.shape), tensor)


def allgather(tensor, name=None):
    """Concatenate ``tensor`` of every worker along dimension 0 (first dimensions may differ)."""
    a = _to_numpy(tensor)
    a = a.reshape(1) if a.ndim == 0 else a
    return _like(_hvd.allgather(torch.from_numpy(a), name=name).numpy(), tensor)


def broadcast(tensor, root_rank, name=None):
    """Value of ``tensor`` on ``root_rank``, on every worker."""
    a = _to_numpy(tensor)
    out = _hvd.broadcast(_torch(a.copy()), root_rank, name=name)
    return _like(out.numpy().reshape(a.shape), tensor)


def broadcast_variables(variables, root_rank=0):
    """Assign every variable (objects with ``numpy()`` + ``assign()``, e.g. ``tf.Variable``) its value on ``root_rank``."""
    for i, v in enumerate(variables):
        new = broadcast(_to_numpy(v), root_rank, name="broadcast_variables.%d" % i)
        v.assign(np.asarray(new).reshape(_to_numpy(v).shape))


def broadcast_object(obj, root_rank=0, name=None):
    """cloudpickle ``obj`` on ``root_rank`` and return it on every rank (``hvd.broadcast_object``)."""
    return _hvd.broadcast_object(obj, root_rank, name)


def allgather_object(obj, name=None):
    """List with every worker's (cloudpickled) ``obj``, in rank order."""
    return _hvd.allgather_object(obj, name=name)


def grouped_allreduce(tensors, average=None, name=None, op=None):
    """Allreduce a list of tensors as one negotiated group (the engine fuses them into one buffer)."""
    arrs = [_to_numpy(t) for t in tensors]
    outs = _hvd.grouped_allreduce([_torch(a) for a in arrs], average=average, name=name, op=op)
    return [_like(o.numpy().reshape(a.shape), t) for o, a, t in zip(outs, arrs, tensors)]


class DistributedGradientTape(object):
    """``hvd.DistributedGradientTape(tape)``: wraps an object with ``gradient(target, sources, ...)`` (a
    ``tf.GradientTape``) and averages the returned gradients over all workers - one grouped allreduce per call, ``None``
    gradients stay ``None``.  Usable as a context manager when the wrapped tape is one."""

    def __init__(self, gradtape, op=None, name=None):
        self._tape = gradtape
        self._op = Average if op is None else op
        self._name = name or "DistributedGradientTape"
        self._calls = 0

    def __enter__(self):
        self._tape.__enter__()
        return self

    def __exit__(self, *exc):
        return self._tape.__exit__(*exc)

    def __getattr__(self, item):                 # watch(), stop_recording(), ... go to the wrapped tape
        return getattr(self._tape, item)

    def gradient(self, target, sources, *args, **kwargs):
        grads = self._tape.gradient(target, sources, *args, **kwargs)
        single = not isinstance(grads, (list, tuple))
        glist = [grads] if single else list(grads)
        idx = [i for i, g in enumerate(glist) if g is not None]
        self._calls += 1
        if idx and size() > 1:
            red = grouped_allreduce([glist[i] for i in idx], op=self._op, name="%s.%d" % (self._name, self._calls))
            for i, g in zip(idx, red):
                glist[i] = g
        return glist[0] if single else (type(grads)(glist) if isinstance(grads, tuple) else glist)
