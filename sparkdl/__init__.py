"""``sparkdl`` - API-compatible facade of the B200-native runtime.

Same public surface as the reference package (``/root/reference/sparkdl/__init__.py:19-24``):
``sparkdl.HorovodRunner`` and ``sparkdl.__version__``.  The implementation lives in
:mod:`spark_deep_learning_b200`; additionally ``sparkdl.hvd`` exposes the Horovod-style
collective frontend that a pickled ``main`` calls.
"""
from sparkdl.horovod.runner_base import HorovodRunner

__all__ = ['HorovodRunner']

__version__ = '2.2.0-db1'
__b200_version__ = '0.1.0'
