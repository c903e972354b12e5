from pyspark.ml import _p

MLReadable, MLWritable = _p.MLReadable, _p.MLWritable
