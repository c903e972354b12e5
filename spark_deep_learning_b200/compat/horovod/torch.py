"""``import horovod.torch as hvd`` -> :mod:`spark_deep_learning_b200.parallel.hvd`."""
import sys

from spark_deep_learning_b200.parallel import hvd as _impl

sys.modules[__name__] = _impl
