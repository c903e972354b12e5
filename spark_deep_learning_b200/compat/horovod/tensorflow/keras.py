"""``horovod.tensorflow.keras``-style API: ``DistributedOptimizer`` and the callbacks of a Keras Horovod main
(``hvd.callbacks.BroadcastGlobalVariablesCallback(0)``, ``MetricAverageCallback()``, ``LearningRateWarmupCallback``),
on top of :mod:`horovod.tensorflow` of this shim.  Works with any optimizer / model that follows the Keras protocol
(``apply_gradients(grads_and_vars)``, ``get_config`` / ``from_config``, ``model.get_weights`` / ``set_weights``), which is
how it is tested on an image without TensorFlow; with TensorFlow installed the same code drives real Keras objects.
"""
import numpy as np

from spark_deep_learning_b200.compat.horovod import tensorflow as _hvd_tf
from spark_deep_learning_b200.compat.horovod.tensorflow import (Average, Sum, allgather, allreduce, barrier,  # noqa: F401
                                                                broadcast, broadcast_object, broadcast_variables, init,
                                                                is_initialized, local_rank, local_size, rank, shutdown,
                                                                size)


def _callback_base():
    try:
        from tensorflow import keras  # pylint: disable=import-error
        return keras.callbacks.Callback
    except Exception:  # pylint: disable=broad-except
        try:
            import keras  # pylint: disable=import-error
            return keras.callbacks.Callback
        except Exception:  # pylint: disable=broad-except
            return object


def DistributedOptimizer(optimizer, name=None, average=True, backward_passes_per_step=1):  # pylint: disable=invalid-name
    """Wrap a Keras optimizer so that ``apply_gradients`` first averages (or sums) every gradient over all workers.

    Like Horovod's, the result is an instance of a dynamically created subclass of ``optimizer``'s own class, rebuilt
    from ``optimizer.get_config()`` when the class supports it (so Keras' serialisation keeps working)."""
    base = optimizer.__class__
    passes = max(1, int(backward_passes_per_step))

    def apply_gradients(self, grads_and_vars, *args, **kwargs):
        gv = list(grads_and_vars)
        self._sparkdl_pass = getattr(self, "_sparkdl_pass", 0) + 1
        if passes > 1:
            acc = getattr(self, "_sparkdl_acc", None)
            if acc is None:
                acc = self._sparkdl_acc = [None] * len(gv)
            for i, (g, _) in enumerate(gv):
                if g is not None:
                    a = _hvd_tf._to_numpy(g)
                    acc[i] = a.copy() if acc[i] is None else acc[i] + a
            if self._sparkdl_pass % passes != 0:
                return None
            gv = [(None if acc[i] is None else acc[i] / passes, v) for i, (_, v) in enumerate(gv)]
            self._sparkdl_acc = None
        reduced = []
        for i, (g, v) in enumerate(gv):
            if g is None:
                reduced.append((g, v))
                continue
            r = _hvd_tf.allreduce(g, average=average, name="%s.grad.%d" % (name or "DistributedOptimizer", i))
            reduced.append((r, v))
        return base.apply_gradients(self, reduced, *args, **kwargs)

    cls = type(base.__name__, (base,), {"apply_gradients": apply_gradients, "_sparkdl_distributed": True})
    if hasattr(optimizer, "get_config") and hasattr(cls, "from_config"):
        try:
            return cls.from_config(optimizer.get_config())
        except Exception:  # pylint: disable=broad-except
            pass
    optimizer.__class__ = cls
    return optimizer


class _Callbacks(object):
    """Namespace mirroring ``horovod.tensorflow.keras.callbacks``."""

    class BroadcastGlobalVariablesCallback(_callback_base()):
        """Broadcast the model (and optimizer) state from ``root_rank`` at the start of training, so that all workers
        start from the same weights (random init or a checkpoint restored on one worker only)."""

        def __init__(self, root_rank=0):
            try:
                super().__init__()
            except TypeError:
                pass
            self.root_rank = root_rank
            self.broadcast_done = False

        def on_batch_end(self, batch, logs=None):
            if self.broadcast_done:
                return
            model = getattr(self, "model", None)
            if model is not None:
                ws = [np.asarray(broadcast(w, self.root_rank, name="bgv.%d" % i)) for i, w in enumerate(model.get_weights())]
                model.set_weights(ws)
                opt = getattr(model, "optimizer", None)
                if opt is not None and hasattr(opt, "get_weights") and hasattr(opt, "set_weights"):
                    ow = opt.get_weights()
                    if ow:
                        opt.set_weights([np.asarray(broadcast(w, self.root_rank, name="bgv.opt.%d" % i)) for i, w in enumerate(ow)])
            self.broadcast_done = True

        on_train_batch_end = on_batch_end

    class MetricAverageCallback(_callback_base()):
        """Average the epoch-end metrics over all workers (so that rank 0 logs / checkpoints on global numbers)."""

        def __init__(self):
            try:
                super().__init__()
            except TypeError:
                pass

        def on_epoch_end(self, epoch, logs=None):
            if not logs:
                return
            for k in sorted(logs):
                try:
                    v = float(logs[k])
                except (TypeError, ValueError):
                    continue
                logs[k] = float(allreduce(np.array([v], dtype=np.float64), average=True, name="metric.%s.%d" % (k, epoch))[0])

    class LearningRateWarmupCallback(_callback_base()):
        """Ramp the learning rate from ``initial_lr / size`` to ``initial_lr`` over the first ``warmup_epochs`` epochs."""

        def __init__(self, initial_lr, warmup_epochs=5, steps_per_epoch=None, verbose=0):
            try:
                super().__init__()
            except TypeError:
                pass
            self.initial_lr, self.warmup_epochs = float(initial_lr), warmup_epochs
            self.steps_per_epoch, self.verbose = steps_per_epoch, verbose
            self._epoch = 0

        def _set_lr(self, lr):
            opt = getattr(getattr(self, "model", None), "optimizer", None)
            if opt is None:
                return
            if hasattr(opt, "learning_rate") and hasattr(opt.learning_rate, "assign"):
                opt.learning_rate.assign(lr)
            else:
                for attr in ("learning_rate", "lr"):
                    if hasattr(opt, attr):
                        setattr(opt, attr, lr)
                        break

        def on_epoch_begin(self, epoch, logs=None):
            self._epoch = epoch
            if epoch < self.warmup_epochs and not self.steps_per_epoch:
                frac = epoch / float(self.warmup_epochs)
                self._set_lr(self.initial_lr * (1.0 / size() * (1 - frac) + frac))
            elif epoch == self.warmup_epochs:
                self._set_lr(self.initial_lr)

        def on_batch_begin(self, batch, logs=None):
            if self.steps_per_epoch and self._epoch < self.warmup_epochs:
                frac = (self._epoch + batch / float(self.steps_per_epoch)) / float(self.warmup_epochs)
                self._set_lr(self.initial_lr * (1.0 / size() * (1 - frac) + frac))

        on_train_batch_begin = on_batch_begin


callbacks = _Callbacks
