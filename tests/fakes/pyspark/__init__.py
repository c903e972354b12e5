"""A tiny stand-in for pyspark (NOT pyspark): just enough API surface to run this repo's Spark code paths without a JVM -
SparkSession / SparkContext.parallelize / RDD.mapPartitions / RDD.barrier() (tasks = local subprocesses started all
together, ``BarrierTaskContext`` with partitionId / barrier / allGather / getTaskInfos), a pandas-backed DataFrame, and
``pyspark.ml`` re-exporting the repo's dependency-free Param machinery.  Test-only."""
from pyspark.context import BarrierTaskContext, SparkContext  # noqa: F401

__version__ = "0.0-fake"
