"""Module path parity with the reference (``sparkdl.xgboost.xgboost``); see
:mod:`spark_deep_learning_b200.xgboost.estimator` for the implementation."""
from spark_deep_learning_b200.xgboost.estimator import (  # noqa: F401
    XgboostClassifier, XgboostClassifierModel, XgboostRegressor, XgboostRegressorModel,
    _XgboostEstimator, _XgboostModel, _XgboostParams)
