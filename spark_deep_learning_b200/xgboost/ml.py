"""Tiny stand-ins for ``pyspark.ml.Pipeline`` / ``CrossValidator`` / ``ParamGridBuilder`` / evaluators over pandas
frames, so the claim "can be used in Pipeline and meta algorithms like CrossValidator"
(``/root/reference/sparkdl/xgboost/xgboost.py:168-169``) is testable without Spark.  With pyspark installed, use
pyspark's own classes - the estimators subclass pyspark's ``Estimator`` / ``Model`` then."""
import itertools

import numpy as np


class Pipeline(object):
    def __init__(self, stages):
        self.stages = list(stages)

    def fit(self, dataset):
        fitted, cur = [], dataset
        for st in self.stages:
            if hasattr(st, "fit"):
                m = st.fit(cur)
                fitted.append(m)
                cur = m.transform(cur)
            else:
                fitted.append(st)
                cur = st.transform(cur)
        return PipelineModel(fitted)


class PipelineModel(object):
    def __init__(self, stages):
        self.stages = stages

    def transform(self, dataset):
        for st in self.stages:
            dataset = st.transform(dataset)
        return dataset


class ParamGridBuilder(object):
    def __init__(self):
        self._grid = {}

    def addGrid(self, param, values):
        self._grid[param] = list(values)
        return self

    def build(self):
        keys = list(self._grid)
        return [dict(zip(keys, combo)) for combo in itertools.product(*[self._grid[k] for k in keys])]


class RegressionEvaluator(object):
    def __init__(self, predictionCol="prediction", labelCol="label", metricName="rmse"):
        self.p, self.l, self.m = predictionCol, labelCol, metricName

    def evaluate(self, df):
        e = np.asarray(df[self.p], dtype=float) - np.asarray(df[self.l], dtype=float)
        return float(np.sqrt(np.mean(e ** 2))) if self.m == "rmse" else float(np.mean(np.abs(e)))

    def isLargerBetter(self):
        return False


class MulticlassClassificationEvaluator(object):
    def __init__(self, predictionCol="prediction", labelCol="label", metricName="accuracy"):
        self.p, self.l = predictionCol, labelCol

    def evaluate(self, df):
        return float(np.mean(np.asarray(df[self.p]) == np.asarray(df[self.l])))

    def isLargerBetter(self):
        return True


class CrossValidator(object):
    def __init__(self, estimator, estimatorParamMaps, evaluator, numFolds=3, seed=0):
        self.estimator, self.maps, self.evaluator, self.k, self.seed = estimator, estimatorParamMaps, evaluator, numFolds, seed

    def fit(self, df):
        n = len(df)
        folds = np.random.RandomState(self.seed).permutation(n) % self.k
        scores = np.zeros(len(self.maps))
        for f in range(self.k):
            tr, va = df[folds != f].reset_index(drop=True), df[folds == f].reset_index(drop=True)
            for i, pm in enumerate(self.maps):
                scores[i] += self.evaluator.evaluate(self.estimator.fit(tr, pm).transform(va)) / self.k
        best = int(np.argmax(scores) if self.evaluator.isLargerBetter() else np.argmin(scores))
        model = CrossValidatorModel(self.estimator.fit(df, self.maps[best]), scores.tolist(), best)
        return model


class CrossValidatorModel(object):
    def __init__(self, bestModel, avgMetrics, bestIndex):
        self.bestModel, self.avgMetrics, self.bestIndex = bestModel, avgMetrics, bestIndex

    def transform(self, df):
        return self.bestModel.transform(df)
