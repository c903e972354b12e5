"""Tiny stand-ins for ``pyspark.ml.Pipeline`` / ``CrossValidator`` / ``TrainValidationSplit`` / ``OneVsRest`` /
``ParamGridBuilder`` / evaluators over pandas frames, so the claim "can be used in Pipeline and meta algorithms like
CrossValidator / TrainValidationSplit / OneVsRest" (``/root/reference/sparkdl/xgboost/xgboost.py:168-169,250-251``)
is testable without Spark.  With pyspark installed, use
pyspark's own classes - the estimators subclass pyspark's ``Estimator`` / ``Model`` then."""
import itertools

import numpy as np


class Pipeline(object):
    """Chain of estimators / transformers; ``fit`` returns a ``PipelineModel``."""
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
    """Fitted pipeline: applies every stage's ``transform`` in order."""
    def __init__(self, stages):
        self.stages = stages

    def transform(self, dataset):
        for st in self.stages:
            dataset = st.transform(dataset)
        return dataset


class ParamGridBuilder(object):
    """Cartesian product of ``addGrid(param, values)`` as a list of param maps."""
    def __init__(self):
        self._grid = {}

    def addGrid(self, param, values):
        self._grid[param] = list(values)
        return self

    def build(self):
        keys = list(self._grid)
        return [dict(zip(keys, combo)) for combo in itertools.product(*[self._grid[k] for k in keys])]


class RegressionEvaluator(object):
    """``rmse`` (default) or ``mae`` between ``predictionCol`` and ``labelCol``."""
    def __init__(self, predictionCol="prediction", labelCol="label", metricName="rmse"):
        self.p, self.l, self.m = predictionCol, labelCol, metricName

    def evaluate(self, df):
        e = np.asarray(df[self.p], dtype=float) - np.asarray(df[self.l], dtype=float)
        return float(np.sqrt(np.mean(e ** 2))) if self.m == "rmse" else float(np.mean(np.abs(e)))

    def isLargerBetter(self):
        return False


class MulticlassClassificationEvaluator(object):
    """Accuracy of ``predictionCol`` against ``labelCol``."""
    def __init__(self, predictionCol="prediction", labelCol="label", metricName="accuracy"):
        self.p, self.l = predictionCol, labelCol

    def evaluate(self, df):
        return float(np.mean(np.asarray(df[self.p]) == np.asarray(df[self.l])))

    def isLargerBetter(self):
        return True


class CrossValidator(object):
    """K-fold model selection over ``estimatorParamMaps``; refits the best map on the whole frame."""
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
    """Result of ``CrossValidator.fit``: ``bestModel``, ``avgMetrics``, ``bestIndex``."""
    def __init__(self, bestModel, avgMetrics, bestIndex):
        self.bestModel, self.avgMetrics, self.bestIndex = bestModel, avgMetrics, bestIndex

    def transform(self, df):
        return self.bestModel.transform(df)


class BinaryClassificationEvaluator(object):
    """areaUnderROC of ``rawPredictionCol`` (margin of the positive class), rank-based with tie averaging."""

    def __init__(self, rawPredictionCol="rawPrediction", labelCol="label", metricName="areaUnderROC"):
        if metricName != "areaUnderROC":
            raise ValueError("only areaUnderROC is implemented")
        self.r, self.l = rawPredictionCol, labelCol

    def evaluate(self, df):
        raw = np.asarray([np.asarray(v, dtype=float).ravel()[-1] for v in df[self.r]])
        y = np.asarray(df[self.l], dtype=float) > 0.5
        npos, nneg = int(y.sum()), int((~y).sum())
        if npos == 0 or nneg == 0:
            return 0.5
        order = np.argsort(raw, kind="mergesort")
        ranks = np.empty(len(raw))
        sr = raw[order]
        i = 0
        while i < len(sr):                                   # average ranks over ties
            j = i
            while j + 1 < len(sr) and sr[j + 1] == sr[i]:
                j += 1
            ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
            i = j + 1
        return float((ranks[y].sum() - npos * (npos + 1) / 2.0) / (npos * nneg))

    def isLargerBetter(self):
        return True


class TrainValidationSplit(object):
    """One random train/validation split per param map (``pyspark.ml.tuning.TrainValidationSplit``)."""

    def __init__(self, estimator, estimatorParamMaps, evaluator, trainRatio=0.75, seed=0):
        if not 0.0 < trainRatio < 1.0:
            raise ValueError("trainRatio must be in (0, 1)")
        self.estimator, self.maps, self.evaluator, self.ratio, self.seed = estimator, estimatorParamMaps, evaluator, trainRatio, seed

    def fit(self, df):
        is_tr = np.random.RandomState(self.seed).rand(len(df)) < self.ratio
        tr, va = df[is_tr].reset_index(drop=True), df[~is_tr].reset_index(drop=True)
        metrics = [self.evaluator.evaluate(self.estimator.fit(tr, pm).transform(va)) for pm in self.maps]
        best = int(np.argmax(metrics) if self.evaluator.isLargerBetter() else np.argmin(metrics))
        return TrainValidationSplitModel(self.estimator.fit(df, self.maps[best]), metrics, best)


class TrainValidationSplitModel(object):
    """Result of ``TrainValidationSplit.fit``: ``bestModel``, ``validationMetrics``, ``bestIndex``."""
    def __init__(self, bestModel, validationMetrics, bestIndex):
        self.bestModel, self.validationMetrics, self.bestIndex = bestModel, validationMetrics, bestIndex

    def transform(self, df):
        return self.bestModel.transform(df)


class OneVsRest(object):
    """Reduce a K-class problem to K binary fits of ``classifier`` (``pyspark.ml.classification.OneVsRest``): class k's
    model is trained on ``label == k``; prediction = argmax over the positive-class margins."""

    def __init__(self, classifier, featuresCol="features", labelCol="label", predictionCol="prediction",
                 rawPredictionCol="rawPrediction"):
        self.classifier = classifier
        self.featuresCol, self.labelCol, self.predictionCol, self.rawPredictionCol = featuresCol, labelCol, predictionCol, rawPredictionCol

    def fit(self, df):
        y = np.asarray(df[self.labelCol], dtype=float)
        k = int(np.nanmax(y)) + 1
        tmp = "__ovr_label__"
        models = []
        for c in range(k):
            d = df.copy()
            d[tmp] = (y == c).astype(np.float32)
            models.append(self.classifier.copy().setParams(labelCol=tmp, featuresCol=self.featuresCol).fit(d))
        return OneVsRestModel(models, self.predictionCol, self.rawPredictionCol)


class OneVsRestModel(object):
    """K binary models; ``transform`` adds the per-class margins and their argmax."""
    def __init__(self, models, predictionCol="prediction", rawPredictionCol="rawPrediction"):
        self.models, self.predictionCol, self.rawPredictionCol = models, predictionCol, rawPredictionCol

    def transform(self, df):
        margins = []
        for m in self.models:
            raw_col = m.getOrDefault("rawPredictionCol")
            out = m.transform(df)
            margins.append(np.asarray([np.asarray(v, dtype=float).ravel()[-1] for v in out[raw_col]]))
        raw = np.stack(margins, axis=1)
        res = df.copy()
        res[self.rawPredictionCol] = list(raw)
        res[self.predictionCol] = raw.argmax(axis=1).astype(float)
        return res
