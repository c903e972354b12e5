"""NVTX ranges around the runtime's kernel launches, enabled with ``SPARKDL_NVTX=1``.

The switch is read once at import: when it is off, :func:`annotate` returns the function object
unchanged, so the hot path carries no wrapper at all.  Ranges are host-side markers (legal during
CUDA-graph capture); view them in Nsight Systems / Nsight Compute next to the kernels they enclose.
"""
import functools
import os

ENABLED = os.environ.get("SPARKDL_NVTX", "0") == "1"


def _push_pop():
    import torch
    return torch.cuda.nvtx.range_push, torch.cuda.nvtx.range_pop


def annotate(name, enabled=None):
    """Decorator: run the function inside an NVTX range called ``name``."""
    on = ENABLED if enabled is None else enabled

    def deco(fn):
        if not on:
            return fn
        push, pop = _push_pop()

        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            push(name)
            try:
                return fn(*args, **kwargs)
            finally:
                pop()
        return wrapper
    return deco


class range(object):  # pylint: disable=redefined-builtin
    """``with nvtx.range("name"):`` - no-op unless ``SPARKDL_NVTX=1``."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if ENABLED:
            _push_pop()[0](self.name)
        return self

    def __exit__(self, *exc):
        if ENABLED:
            _push_pop()[1]()
        return False
