"""``import sparkdl.hvd as hvd`` - the Horovod-style frontend (alias of
:mod:`spark_deep_learning_b200.parallel.hvd`; one module object, one runtime state)."""
import sys

from spark_deep_learning_b200.parallel import hvd as _impl

sys.modules[__name__] = _impl
