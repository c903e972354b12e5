"""Torch-side companion of ``sparkdl.horovod.tensorflow.keras.LogCallback`` (SURVEY.md §5.5:
"add a framework-neutral callback + a torch-friendly helper alongside the Keras one")."""
from sparkdl.horovod.tensorflow.keras import LogCallback as _KerasStyle

__all__ = ["LogCallback"]


class LogCallback(_KerasStyle):
    """Use from a plain PyTorch loop on rank 0::

        cb = LogCallback(per_batch_log=True)
        for epoch in range(E):
            cb.on_epoch_begin(epoch)
            for i, batch in enumerate(loader):
                ...
                cb.on_batch_end(i, {"loss": loss.item()})
            cb.on_epoch_end(epoch, {"loss": epoch_loss})
    """
