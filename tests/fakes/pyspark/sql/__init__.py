import re

import numpy as np

from pyspark.context import RDD, SparkContext


class Row(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def asDict(self):
        return dict(self)


class DataFrame(object):
    """pandas-backed; partitions are contiguous row ranges."""

    def __init__(self, session, pdf, nparts=None):
        self.sparkSession = session
        self._pdf = pdf.reset_index(drop=True)
        self._n = nparts or min(len(pdf), session.sparkContext.defaultParallelism) or 1

    @property
    def columns(self):
        return list(self._pdf.columns)

    def toPandas(self):
        return self._pdf.copy()

    def count(self):
        return len(self._pdf)

    def repartition(self, n):
        perm = np.random.RandomState(0).permutation(len(self._pdf))
        return DataFrame(self.sparkSession, self._pdf.iloc[perm], n)

    def select(self, *cols):
        return DataFrame(self.sparkSession, self._pdf[list(cols)], self._n)

    @property
    def rdd(self):
        recs = [Row(r) for r in self._pdf.to_dict("records")]
        parts = [list(p) for p in np.array_split(np.arange(len(recs)), self._n)]
        return RDD(self.sparkSession.sparkContext, [[recs[i] for i in p] for p in parts])


class _Builder(object):
    def __init__(self):
        self._master = "local[4]"

    def master(self, m):
        self._master = m
        return self

    def appName(self, _):
        return self

    def config(self, *a, **k):
        return self

    def getOrCreate(self):
        if SparkSession._active is None:
            m = re.match(r"local\[(\d+)\]", self._master)
            SparkSession._active = SparkSession(int(m.group(1)) if m else 4)
        return SparkSession._active


class SparkSession(object):
    _active = None
    builder = _Builder()

    def __init__(self, slots):
        self.sparkContext = SparkContext(slots)

    @classmethod
    def getActiveSession(cls):
        return cls._active

    def createDataFrame(self, pdf):
        return DataFrame(self, pdf)

    def stop(self):
        SparkSession._active = None
