from pyspark.ml import _p


def _with_default(base):
    """pyspark's shared mixins set their default in __init__ (the repo's shim applies it separately)."""
    name, default = base.__dict__["_shared_default"]

    class K(base):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            if default is not None:
                self._setDefault(**{name: default})
    K.__name__ = K.__qualname__ = base.__name__
    return K


HasFeaturesCol, HasLabelCol, HasPredictionCol = (_with_default(c) for c in (_p.HasFeaturesCol, _p.HasLabelCol, _p.HasPredictionCol))
HasProbabilityCol, HasRawPredictionCol = _with_default(_p.HasProbabilityCol), _with_default(_p.HasRawPredictionCol)
HasValidationIndicatorCol, HasWeightCol = _with_default(_p.HasValidationIndicatorCol), _with_default(_p.HasWeightCol)
