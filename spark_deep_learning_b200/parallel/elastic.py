"""``hvd.elastic``: commit / restore / sync state objects and the ``run`` wrapper.

The documented sparkdl contract is a fixed-size gang ("tasks starting all together",
``/root/reference/sparkdl/horovod/runner_base.py:54``; SURVEY.md §5.3: no elastic mode), so a lost
rank still ends the job - what this module gives is the *state* half of Horovod's elastic API so
training loops written against it run unchanged:

* ``TorchState(model, optimizer, **scalars)`` keeps an in-memory snapshot: ``commit()`` saves,
  ``restore()`` rolls back, ``sync()`` broadcasts rank 0's values to every rank;
* ``run(fn)`` calls ``fn(state, ...)`` after ``state.sync()``; if a collective fails with
  ``HorovodInternalError`` the state is rolled back to the last commit before the error propagates
  (the launcher then tears the gang down and re-raises on the driver), so a driver-side retry
  loop can resume from ``state`` persisted by ``on_reset`` / commit callbacks.
"""
import copy
import functools


class ObjectState(object):
    """Plain-python attributes with commit / restore / sync."""

    def __init__(self, **kwargs):
        self._saved = {}
        self._reset_callbacks = []
        self._commit_callbacks = []
        for k, v in kwargs.items():
            setattr(self, k, v)
        self._keys = list(kwargs)
        self.save()

    # -- callbacks ----------------------------------------------------------------------------
    def register_reset_callbacks(self, callbacks):
        self._reset_callbacks.extend(callbacks)

    def register_commit_callbacks(self, callbacks):
        self._commit_callbacks.extend(callbacks)

    def on_reset(self):
        for cb in self._reset_callbacks:
            cb()

    # -- snapshot -----------------------------------------------------------------------------
    def save(self):
        self._saved = {k: copy.deepcopy(getattr(self, k)) for k in self._keys}

    def commit(self):
        self.save()
        for cb in self._commit_callbacks:
            cb(self)

    def restore(self):
        for k, v in self._saved.items():
            setattr(self, k, copy.deepcopy(v))

    def sync(self):
        from spark_deep_learning_b200.parallel import hvd
        if self._keys and hvd.size() > 1:
            vals = hvd.broadcast_object({k: getattr(self, k) for k in self._keys}, root_rank=0, name="elastic.state")
            for k, v in vals.items():
                setattr(self, k, v)
        self.save()


class TorchState(ObjectState):
    """``ObjectState`` plus a model and an optimizer (``state_dict`` snapshots held on the host)."""

    def __init__(self, model=None, optimizer=None, **kwargs):
        self.model, self.optimizer = model, optimizer
        self._model_saved = self._opt_saved = None
        super().__init__(**kwargs)

    @staticmethod
    def _host_copy(sd):
        import torch

        def conv(v):
            if isinstance(v, torch.Tensor):
                return v.detach().to("cpu", copy=True)
            if isinstance(v, dict):
                return {k: conv(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(conv(x) for x in v)
            return copy.deepcopy(v)
        return conv(sd)

    def save(self):
        super().save()
        if self.model is not None:
            self._model_saved = self._host_copy(self.model.state_dict())
        if self.optimizer is not None:
            self._opt_saved = self._host_copy(self.optimizer.state_dict())

    def restore(self):
        super().restore()
        if self.model is not None and self._model_saved is not None:
            self.model.load_state_dict(self._model_saved)
        if self.optimizer is not None and self._opt_saved is not None:
            self.optimizer.load_state_dict(self._opt_saved)

    def sync(self):
        from spark_deep_learning_b200.parallel import hvd
        if hvd.size() > 1:
            if self.model is not None:
                hvd.broadcast_parameters(self.model.state_dict(), root_rank=0)
            if self.optimizer is not None:
                hvd.broadcast_optimizer_state(self.optimizer, root_rank=0)
        super().sync()


def run(func):
    """Decorator: ``func(state, *args, **kwargs)`` with sync-before and rollback-on-collective-failure."""

    @functools.wraps(func)
    def wrapper(state, *args, **kwargs):
        from spark_deep_learning_b200.parallel import hvd
        state.sync()
        try:
            return func(state, *args, **kwargs)
        except hvd.HorovodInternalError:
            state.restore()
            state.on_reset()
            raise
    return wrapper
