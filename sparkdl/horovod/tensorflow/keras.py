"""``LogCallback`` - streams Keras training events to the driver (notebook cell output).

The reference class is a stub whose four methods raise
(``/root/reference/sparkdl/horovod/tensorflow/keras.py:16-34``).  Implemented here: epoch begin
records the wall-clock start, every epoch end (and, with ``per_batch_log=True``, every batch end)
formats the ``logs`` dict and sends it through :func:`sparkdl.horovod.log_to_driver`.

TensorFlow is not installed on the B200 image, so ``keras`` is imported lazily: with TF/Keras
present ``LogCallback`` subclasses ``keras.callbacks.Callback`` exactly like the reference; without
it the same class works as a duck-typed callback (and is what the torch helper below builds on).
"""
import time

from sparkdl.horovod import log_to_driver

__all__ = ["LogCallback"]


def _keras_callback_base():
    try:
        from tensorflow import keras  # pylint: disable=import-error
        return keras.callbacks.Callback
    except Exception:  # pylint: disable=broad-except
        try:
            import keras  # pylint: disable=import-error
            return keras.callbacks.Callback
        except Exception:  # pylint: disable=broad-except
            return object


def _fmt(logs):
    if not logs:
        return ""
    parts = []
    for k in sorted(logs):
        v = logs[k]
        try:
            parts.append("%s: %.4f" % (k, float(v)))
        except (TypeError, ValueError):
            parts.append("%s: %s" % (k, v))
    return " - ".join(parts)


class LogCallback(_keras_callback_base()):
    """
    A simple HorovodRunner log callback that streams event logs to notebook cell output.
    """

    def __init__(self, per_batch_log=False):
        """
        :param per_batch_log: whether to output logs per batch, default: False.
        """
        try:
            super().__init__()
        except TypeError:
            pass
        self.per_batch_log = per_batch_log
        self._epoch = 0
        self._epoch_start = None

    def on_epoch_begin(self, epoch, logs=None):
        self._epoch = epoch
        self._epoch_start = time.time()
        log_to_driver("Epoch %d begin at %s" % (epoch, time.strftime("%Y-%m-%d %H:%M:%S")))

    def on_batch_end(self, batch, logs=None):
        if self.per_batch_log:
            log_to_driver("Epoch %d batch %d: %s" % (self._epoch, batch, _fmt(logs)))

    def on_epoch_end(self, epoch, logs=None):
        elapsed = time.time() - self._epoch_start if self._epoch_start is not None else 0.0
        log_to_driver("Epoch %d end, %.3f seconds elapsed. %s" % (epoch, elapsed, _fmt(logs)))
