"""``sparkdl.xgboost`` - same four public classes as the reference
(``/root/reference/sparkdl/xgboost/__init__.py:19-23``), implemented on the B200-native GBDT engine.
``Vectors`` is exported for convenience on hosts without pyspark."""
from sparkdl.xgboost.xgboost import XgboostClassifier, XgboostClassifierModel, \
    XgboostRegressor, XgboostRegressorModel
from spark_deep_learning_b200.xgboost.linalg import Vectors

__all__ = ['XgboostClassifier', 'XgboostClassifierModel', 'XgboostRegressor', 'XgboostRegressorModel']
