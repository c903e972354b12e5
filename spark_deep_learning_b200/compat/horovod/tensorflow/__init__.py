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
           "broadcast_variables", "broadcast_object", "Average", "Sum", "Min", "Max", "Product", "barrier"]


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
