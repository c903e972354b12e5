from pyspark.ml import _p

Param, Params, TypeConverters = _p.Param, _p.Params, _p.TypeConverters
