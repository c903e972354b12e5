"""From-scratch histogram GBDT + Spark-ML style estimators (the engine under ``sparkdl.xgboost``)."""
from spark_deep_learning_b200.xgboost.estimator import (XgboostClassifier, XgboostClassifierModel,  # noqa: F401
                                                        XgboostRegressor, XgboostRegressorModel)
from spark_deep_learning_b200.xgboost.gbdt import Booster, train  # noqa: F401
from spark_deep_learning_b200.xgboost.linalg import Vectors  # noqa: F401
