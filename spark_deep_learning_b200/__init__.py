"""spark_deep_learning_b200 - B200-native runtime behind the ``sparkdl`` API.

Layout:

* ``parallel/``  launcher, rendezvous store, C++ engine binding, hvd frontend, DistributedOptimizer
* ``ops/``       ctypes bindings of the sm_100a kernels (collectives, fused optimizer, norm, GEMM, GBDT)
* ``models/``    ResNet-50 / BERT-large built on those ops
* ``xgboost/``   histogram GBDT trainer + Spark-ML style estimators
* ``utils/``     config, timeline, clocks sampler
* ``csrc/``      C++ (core) and CUDA (cuda) sources, built in-tree by :mod:`build`
"""
__version__ = "0.1.0"
