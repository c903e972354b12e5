"""Compatibility shims.  ``compat/horovod`` is an importable ``horovod`` package whose ``horovod.torch``
is our frontend, for unmodified Horovod training scripts ("an importable horovod.torch alias
shim only if horovod is not installed").  Opt in with ``SPARKDL_HOROVOD_SHIM=1`` (the launcher then puts this
directory on the workers' ``PYTHONPATH``) or call :func:`install_horovod_shim`."""
import os
import sys


def install_horovod_shim():
    try:
        import horovod  # noqa: F401  (a real Horovod wins)
        if not getattr(horovod, "__sparkdl_shim__", False):
            return False
    except ImportError:
        pass
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import horovod  # noqa: F401
    return True
