"""Minimal ``Vectors.dense`` / ``Vectors.sparse`` so the reference doctests
(``/root/reference/sparkdl/xgboost/xgboost.py:223-239``) can be written against pandas frames when
pyspark is absent.  With pyspark installed, ``pyspark.ml.linalg`` vectors are accepted as well
(anything with ``toArray()``)."""
import numpy as np


class DenseVector(object):
    def __init__(self, values):
        self.values = np.asarray(values, dtype=np.float64).reshape(-1)

    def toArray(self):
        return self.values

    def __len__(self):
        return self.values.shape[0]

    def __repr__(self):
        return "DenseVector(%s)" % (self.values.tolist(),)


class SparseVector(object):
    def __init__(self, size, *args):
        self.size = int(size)
        if len(args) == 1:
            pairs = args[0]
            if isinstance(pairs, dict):
                pairs = sorted(pairs.items())
            self.indices = np.asarray([p[0] for p in pairs], dtype=np.int64)
            self.values = np.asarray([p[1] for p in pairs], dtype=np.float64)
        else:
            self.indices = np.asarray(args[0], dtype=np.int64)
            self.values = np.asarray(args[1], dtype=np.float64)

    def toArray(self):
        a = np.zeros(self.size, dtype=np.float64)   # inactive entries mean 0 (ref xgboost.py:46-47)
        a[self.indices] = self.values
        return a

    def __len__(self):
        return self.size

    def __repr__(self):
        return "SparseVector(%d, %s)" % (self.size, dict(zip(self.indices.tolist(), self.values.tolist())))


class Vectors(object):
    @staticmethod
    def dense(*values):
        if len(values) == 1 and hasattr(values[0], "__iter__"):
            values = values[0]
        return DenseVector(values)

    @staticmethod
    def sparse(size, *args):
        return SparseVector(size, *args)


def features_to_matrix(column):
    """A column of vectors / lists / arrays -> float32 matrix [n, d]."""
    if isinstance(column, np.ndarray) and column.ndim == 2:
        return np.ascontiguousarray(column, dtype=np.float32)
    rows = []
    for v in column:
        if hasattr(v, "toArray"):
            v = v.toArray()
        rows.append(np.asarray(v, dtype=np.float32).reshape(-1))
    if not rows:
        return np.zeros((0, 0), dtype=np.float32)
    return np.ascontiguousarray(np.stack(rows))
