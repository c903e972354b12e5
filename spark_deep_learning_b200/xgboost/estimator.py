"""``XgboostRegressor`` / ``XgboostClassifier`` and their models - Spark-ML style estimators over the
from-scratch GBDT engine.

API parity with ``/root/reference/sparkdl/xgboost/xgboost.py``: same class graph (``:38-163``), the
same eight special params (``:41-106``), ``**kwargs`` constructors that accept the
``xgboost.XGBRegressor/XGBClassifier`` constructor + fit/predict parameters (``:171-174``), the same
list of unsupported arguments with the same redirections (``:176-197``), ``rawPredictionCol`` /
``probabilityCol`` on the classifier (``:264-276``), ``get_booster`` (``:130-133``) and
``MLReadable/MLWritable`` persistence (``:117-122,136-141``).  Every method that raises
``NotImplementedError`` in the reference is implemented here.

Datasets: pandas DataFrames (also dicts of columns, and pyspark DataFrames when pyspark is installed).
The features column holds ``Vectors.dense/sparse`` (see :mod:`.linalg`), lists or arrays.
"""
import os
import tempfile

import numpy as np

from spark_deep_learning_b200.xgboost import gbdt
from spark_deep_learning_b200.xgboost.linalg import features_to_matrix
from spark_deep_learning_b200.xgboost.params import (
    HAVE_PYSPARK, Estimator, HasFeaturesCol, HasLabelCol, HasPredictionCol, HasProbabilityCol,
    HasRawPredictionCol, HasValidationIndicatorCol, HasWeightCol, MLReadable, MLWritable, Model, Param,
    Params, TypeConverters, _apply_shared_defaults, load_metadata, save_metadata)

# xgboost.XGBModel constructor / fit / predict parameters that become Params automatically
_SKLEARN_PARAMS = {
    "n_estimators": 100, "max_depth": 6, "learning_rate": 0.3, "verbosity": None, "objective": None,
    "booster": "gbtree", "tree_method": "hist", "n_jobs": None, "gamma": 0.0, "min_child_weight": 1.0,
    "max_delta_step": 0.0, "subsample": 1.0, "colsample_bytree": 1.0, "colsample_bylevel": 1.0,
    "colsample_bynode": 1.0, "reg_alpha": 0.0, "reg_lambda": 1.0, "scale_pos_weight": 1.0, "base_score": 0.5,
    "random_state": 0, "num_parallel_tree": 1, "monotone_constraints": None, "interaction_constraints": None,
    "importance_type": "gain", "max_bin": 256, "eval_metric": None, "early_stopping_rounds": None,
    "verbose": False, "xgb_model": None, "ntree_limit": None, "num_class": None, "tweedie_variance_power": 1.5,
}
_UNSUPPORTED = {
    "gpu_id": "use the `use_gpu` param instead",
    "base_margin": "set the `baseMarginCol` param instead",
    "base_margin_eval_set": "set the `baseMarginCol` param instead",
    "validate_features": "not supported",
    "output_margin": "not supported (the classifier always returns margins in rawPredictionCol)",
    "eval_set": "set the `validationIndicatorCol` param instead",
    "sample_weight": "set the `weightCol` param instead",
    "sample_weight_eval_set": "set the `weightCol` param instead",
}


def _mk(name, doc, conv=None):
    return Param(Params._dummy(), name, doc, typeConverter=conv)


class _XgboostParams(HasFeaturesCol, HasLabelCol, HasWeightCol, HasPredictionCol, HasValidationIndicatorCol):
    missing = _mk("missing", "Specify the missing value in the features, default np.nan. We recommend using 0.0 as the "
                  "missing value for better performance. Note: inactive values in a sparse vector mean 0 instead of "
                  "missing values, unless missing=0 is specified.")
    callbacks = _mk("callbacks", "Callbacks invoked after every boosting round (callable(env) -> stop?). Saved using "
                    "cloudpickle, which is not a fully self-contained format.")
    num_workers = _mk("num_workers", "The number of XGBoost workers. Each worker is one task / one GPU.",
                      TypeConverters.toInt)
    use_gpu = _mk("use_gpu", "A boolean variable. Set use_gpu=true to train on GPUs; one GPU per worker.")
    force_repartition = _mk("force_repartition", "A boolean variable. Force the input dataset to be re-partitioned "
                            "(shuffled into equal shards) before distributed training.")
    use_external_storage = _mk("use_external_storage", "A boolean variable (False by default). Spill the feature matrix "
                               "to disk. Base margin and weighting do not work if this is True.")
    external_storage_precision = _mk("external_storage_precision", "Significant digits kept when spilling to disk.",
                                     TypeConverters.toInt)
    baseMarginCol = _mk("baseMarginCol", "Column with base margins of the training and validation rows. Not available "
                        "for distributed training.")


for _n in _SKLEARN_PARAMS:
    setattr(_XgboostParams, _n, _mk(_n, "xgboost parameter `%s` (see the xgboost scikit-learn API docs)." % _n))


# ---- dataset adapters ---------------------------------------------------------------------------------
class _Frame(object):
    def __init__(self, dataset):
        self.kind = "pandas"
        self.spark = None
        if isinstance(dataset, dict):
            import pandas as pd
            self.kind = "dict"
            self.df = pd.DataFrame({k: list(v) for k, v in dataset.items()})
        elif hasattr(dataset, "toPandas") and hasattr(dataset, "sparkSession"):
            self.kind = "spark"
            self.spark = dataset.sparkSession
            self.df = dataset.toPandas()
        else:
            self.df = dataset

    def has(self, col):
        return col is not None and col != "" and col in self.df.columns

    def col(self, col):
        return self.df[col].to_numpy() if hasattr(self.df[col], "to_numpy") else np.asarray(self.df[col])

    def with_columns(self, cols):
        out = self.df.copy()
        for k, v in cols.items():
            out[k] = list(v) if (isinstance(v, np.ndarray) and v.ndim > 1) else v
        if self.kind == "dict":
            return {k: out[k].tolist() for k in out.columns}
        if self.kind == "spark":  # pragma: no cover
            return self.spark.createDataFrame(out)
        return out


def _json_safe(v):
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    if isinstance(v, float) and v != v:
        return "NaN"
    return v


class _XgboostEstimator(Estimator, _XgboostParams, MLReadable, MLWritable):
    _objective_default = "reg:squarederror"
    _is_classifier = False

    def __init__(self, **kwargs):
        super().__init__()
        _apply_shared_defaults(self)
        self._setDefault(missing=float("nan"), callbacks=None, num_workers=1, use_gpu=False, force_repartition=False,
                         use_external_storage=False, external_storage_precision=5, baseMarginCol=None,
                         validationIndicatorCol=None, weightCol=None, featuresCol="features", labelCol="label",
                         predictionCol="prediction", **_SKLEARN_PARAMS)
        self.setParams(**kwargs)

    def setParams(self, **kwargs):
        for k, v in kwargs.items():
            if k in _UNSUPPORTED:
                raise ValueError("Unsupported param '%s': %s." % (k, _UNSUPPORTED[k]))
            if not self.hasParam(k):
                known = ", ".join(sorted(p.name for p in self.params))
                raise ValueError("Invalid param name: '%s'. Supported params: %s" % (k, known))
            self._set(**{k: v})
        return self

    def _gbdt_params(self):
        p = {k: self.getOrDefault(k) for k in _SKLEARN_PARAMS if k not in ("xgb_model", "ntree_limit")}
        p["use_gpu"] = bool(self.getOrDefault("use_gpu"))
        return p

    def _model_class(self):
        raise NotImplementedError()

    def _fit_on_spark(self, dataset, num_workers):
        """Spark-native training: "Each XGBoost worker corresponds to one spark task" (ref ``xgboost.py:58-64``) - the
        rows never visit the driver.  The (re)partitioned DataFrame runs ONE barrier stage with ``num_workers`` tasks;
        every task turns its partition into a shard, joins the collective gang (``parallel/spark_gang.py``) and trains
        with histogram allreduce; task 0 returns the model."""
        from spark_deep_learning_b200.parallel import spark_gang
        gv = self.getOrDefault
        fcol, lcol, wcol, vcol = gv("featuresCol"), gv("labelCol"), gv("weightCol"), gv("validationIndicatorCol")
        have = set(dataset.columns)
        wcol = wcol if wcol and wcol in have else None
        vcol = vcol if vcol and vcol in have else None
        if gv("baseMarginCol") and gv("baseMarginCol") in have:
            raise ValueError("baseMarginCol is not available for distributed training")
        df = dataset.select(*[c for c in (fcol, lcol, wcol, vcol) if c])
        if bool(gv("force_repartition")) or df.rdd.getNumPartitions() != num_workers:
            df = df.repartition(num_workers)
        params = self._gbdt_params()
        is_clf, default_obj = self._is_classifier, self._objective_default
        extra = dict(missing=float(gv("missing")), xgb_model=gv("xgb_model"), callbacks=gv("callbacks"))
        use_gpu = bool(gv("use_gpu"))

        def task(rows):
            import numpy as _np
            from pyspark import BarrierTaskContext
            from spark_deep_learning_b200.parallel import spark_gang as sg
            from spark_deep_learning_b200.xgboost import gbdt as g
            from spark_deep_learning_b200.xgboost.linalg import features_to_matrix as f2m
            ctx = BarrierTaskContext.get()
            if not use_gpu:
                os.environ["SPARKDL_FORCE_CPU"] = "1"
            rank, server, _lines = sg.enter_gang(ctx, num_workers, tag="xgb")
            out = None
            try:
                rows = list(rows)
                if not rows:
                    raise ValueError("xgboost worker %d received an empty partition; use force_repartition=True" % rank)
                X = f2m([r[fcol] for r in rows])
                y = _np.asarray([r[lcol] for r in rows], dtype=_np.float32)
                w = _np.asarray([r[wcol] for r in rows], dtype=_np.float32) if wcol else None
                val = _np.asarray([bool(r[vcol]) for r in rows]) if vcol else _np.zeros(len(y), dtype=bool)
                import torch
                import sparkdl.hvd as hvd
                hvd.init()
                p = dict(params)
                if is_clf:
                    ymax = float(_np.nanmax(y)) if len(y) else 0.0
                    n_class = int(hvd.allreduce(torch.tensor([ymax]), op=hvd.Max, name="xgb.nclass")[0].item()) + 1
                    if p["objective"] is None:
                        p["objective"] = "binary:logistic" if n_class <= 2 else "multi:softprob"
                    if p["objective"].startswith("multi:"):
                        p["num_class"] = n_class
                elif p["objective"] is None:
                    p["objective"] = default_obj
                any_val = bool(hvd.allreduce(torch.tensor([float(val.any())]), op=hvd.Max, name="xgb.anyval")[0].item())
                tr = ~val
                booster, evals = g.train(params=p, X=X[tr], y=y[tr], weight=None if w is None else w[tr],
                                         eval_set=(X[val], y[val], None if w is None else w[val]) if any_val else None,
                                         base_margin=None, eval_base_margin=None, **extra)
                if rank == 0:
                    out = (booster.to_json(), evals)
            finally:
                sg.leave_gang(ctx, server)
            if rank == 0:
                yield out

        slots = spark_gang.total_slots(dataset.sparkSession.sparkContext)
        if num_workers > slots:
            raise RuntimeError("num_workers=%d needs %d Spark task slots, the cluster has %d" % (num_workers, num_workers, slots))
        js, evals = df.rdd.barrier().mapPartitions(task).collect()[0]
        return gbdt.Booster.from_json(js), evals

    def _finish_model(self, booster, evals):
        model = self._model_class()(xgb_sklearn_model=booster)
        model._evals_result = evals
        for p in self.params:
            if model.hasParam(p.name) and self.isDefined(p):
                model._set(**{p.name: self.getOrDefault(p)}) if self.isSet(p) else None
        return model

    def _fit(self, dataset):
        if (hasattr(dataset, "rdd") and hasattr(dataset, "sparkSession") and int(self.getOrDefault("num_workers") or 1) > 1
                and not self.getOrDefault("use_external_storage")):
            return self._finish_model(*self._fit_on_spark(dataset, int(self.getOrDefault("num_workers"))))
        fr = _Frame(dataset)
        gv = self.getOrDefault
        X = features_to_matrix(fr.col(gv("featuresCol")))
        y = np.asarray(fr.col(gv("labelCol")), dtype=np.float32)
        w = np.asarray(fr.col(gv("weightCol")), dtype=np.float32) if fr.has(gv("weightCol")) else None
        bm = np.asarray(fr.col(gv("baseMarginCol")), dtype=np.float32) if fr.has(gv("baseMarginCol")) else None
        is_val = np.asarray(fr.col(gv("validationIndicatorCol")), dtype=bool) if fr.has(gv("validationIndicatorCol")) \
            else np.zeros(len(y), dtype=bool)
        num_workers = int(gv("num_workers") or 1)
        if gv("use_external_storage"):
            if w is not None or bm is not None:
                raise ValueError("base margin and weighting do not work if use_external_storage is True")
            X = _spill_to_disk(X, int(gv("external_storage_precision") or 5))
        if bm is not None and num_workers > 1:
            raise ValueError("baseMarginCol is not available for distributed training")
        params = self._gbdt_params()
        if self._is_classifier:
            classes = np.unique(y[~np.isnan(y)])
            n_class = int(classes.max()) + 1 if classes.size else 2
            if params["objective"] is None:
                params["objective"] = "binary:logistic" if n_class <= 2 else "multi:softprob"
            if params["objective"].startswith("multi:"):
                params["num_class"] = n_class
        elif params["objective"] is None:
            params["objective"] = self._objective_default
        tr = ~is_val
        train_args = dict(params=params, X=X[tr], y=y[tr], weight=None if w is None else w[tr],
                          eval_set=(X[is_val], y[is_val], None if w is None else w[is_val]) if is_val.any() else None,
                          base_margin=None if bm is None else bm[tr], eval_base_margin=None if bm is None else bm[is_val],
                          missing=float(gv("missing")), xgb_model=gv("xgb_model"), callbacks=gv("callbacks"))
        if num_workers > 1:
            booster, evals = _fit_distributed(train_args, num_workers, bool(gv("force_repartition")),
                                              int(params.get("random_state") or 0))
        else:
            booster, evals = gbdt.train(**train_args)
        return self._finish_model(booster, evals)

    # -- persistence -----------------------------------------------------------------------------------
    def _save_impl(self, path):
        import cloudpickle
        vals = {}
        for p in self.params:
            if self.isSet(p) and p.name not in ("callbacks", "xgb_model"):
                vals[p.name] = _json_safe(self.getOrDefault(p))
        save_metadata(path, type(self).__name__, self.uid, vals)
        if self.isSet("callbacks") and self.getOrDefault("callbacks") is not None:
            with open(os.path.join(path, "callbacks.pkl"), "wb") as f:
                cloudpickle.dump(self.getOrDefault("callbacks"), f)
        xm = self.getOrDefault("xgb_model")
        if isinstance(xm, gbdt.Booster):
            xm.save_model(os.path.join(path, "xgb_model.json"))

    @classmethod
    def _load_impl(cls, path):
        import cloudpickle
        meta = load_metadata(path)
        kw = {k: (float("nan") if v == "NaN" else v) for k, v in meta["paramMap"].items()}
        inst = cls(**kw)
        cb = os.path.join(path, "callbacks.pkl")
        if os.path.exists(cb):
            with open(cb, "rb") as f:
                inst._set(callbacks=cloudpickle.load(f))
        xm = os.path.join(path, "xgb_model.json")
        if os.path.exists(xm):
            inst._set(xgb_model=gbdt.Booster.load_model(xm))
        return inst

    if HAVE_PYSPARK:  # pragma: no cover - pyspark's MLWritable expects a writer object
        def write(self):
            return _SimpleWriter(self)

        @classmethod
        def read(cls):
            return _SimpleReader(cls)


class _XgboostModel(Model, _XgboostParams, MLReadable, MLWritable):
    def __init__(self, xgb_sklearn_model=None):
        super().__init__()
        _apply_shared_defaults(self)
        self._setDefault(missing=float("nan"), featuresCol="features", labelCol="label", predictionCol="prediction",
                         baseMarginCol=None, use_gpu=False, ntree_limit=None)
        self._xgb_sklearn_model = xgb_sklearn_model
        self._evals_result = {}

    def get_booster(self):
        """
        Return the booster instance (:class:`spark_deep_learning_b200.xgboost.gbdt.Booster`).
        """
        return self._xgb_sklearn_model

    def evals_result(self):
        return self._evals_result

    def get_feature_importances(self, importance_type="weight"):
        """``{"f<index>": score}`` as ``Booster.get_score`` (``weight``, ``gain``, ``cover``, ``total_gain``,
        ``total_cover``); the model's ``importance_type`` param is the default of :attr:`feature_importances_`."""
        return self._xgb_sklearn_model.get_score(importance_type=importance_type)

    @property
    def feature_importances_(self):
        """Normalised dense importance vector (``XGBModel.feature_importances_`` semantics)."""
        kind = "gain"
        try:
            kind = self.getOrDefault("importance_type") or "gain"
        except Exception:  # pylint: disable=broad-except
            pass
        return self._xgb_sklearn_model.feature_importances(kind)

    def _margins(self, fr):
        import torch
        gv = self.getOrDefault
        X = features_to_matrix(fr.col(gv("featuresCol")))
        bm = np.asarray(fr.col(gv("baseMarginCol")), dtype=np.float32) if fr.has(gv("baseMarginCol")) else None
        Xt = torch.from_numpy(X)
        if gv("use_gpu") and torch.cuda.is_available():
            Xt = Xt.cuda()
        b = self._xgb_sklearn_model
        limit = gv("ntree_limit")
        if limit is None and b.best_iteration is not None:
            limit = (b.best_iteration + 1) * b.trees_per_round
        return b.predict_margin(Xt, bm, ntree_limit=limit).cpu()

    def _transform(self, dataset):
        fr = _Frame(dataset)
        m = self._margins(fr)
        b = self._xgb_sklearn_model
        pred = gbdt._Objective(b.objective, b.num_class).transform(m)      # inverse link (identity for squared error)
        return fr.with_columns({self.getOrDefault("predictionCol"): pred.numpy().astype(np.float64)})

    def _save_impl(self, path):
        vals = {p.name: _json_safe(self.getOrDefault(p)) for p in self.params
                if self.isSet(p) and p.name not in ("callbacks", "xgb_model")}
        save_metadata(path, type(self).__name__, self.uid, vals)
        self._xgb_sklearn_model.save_model(os.path.join(path, "model.json"))

    @classmethod
    def _load_impl(cls, path):
        meta = load_metadata(path)
        inst = cls(xgb_sklearn_model=gbdt.Booster.load_model(os.path.join(path, "model.json")))
        for k, v in meta["paramMap"].items():
            if inst.hasParam(k):
                inst._set(**{k: float("nan") if v == "NaN" else v})
        return inst

    if HAVE_PYSPARK:  # pragma: no cover
        def write(self):
            return _SimpleWriter(self)

        @classmethod
        def read(cls):
            return _SimpleReader(cls)


class _SimpleWriter(object):  # pragma: no cover - only used with pyspark's MLWritable contract
    def __init__(self, inst):
        self.inst, self._ow = inst, False

    def overwrite(self):
        self._ow = True
        return self

    def save(self, path):
        if os.path.exists(path):
            if not self._ow:
                raise IOError("Path %s already exists." % path)
            import shutil
            shutil.rmtree(path)
        os.makedirs(path)
        self.inst._save_impl(path)


class _SimpleReader(object):  # pragma: no cover
    def __init__(self, cls):
        self.cls = cls

    def load(self, path):
        return self.cls._load_impl(path)


class XgboostRegressorModel(_XgboostModel):
    """
    The model returned by :func:`sparkdl.xgboost.XgboostRegressor.fit`

    .. Note:: This API is experimental.
    """


class XgboostClassifierModel(_XgboostModel, HasProbabilityCol, HasRawPredictionCol):
    """
    The model returned by :func:`sparkdl.xgboost.XgboostClassifier.fit`

    .. Note:: This API is experimental.
    """

    def __init__(self, xgb_sklearn_model=None):
        super().__init__(xgb_sklearn_model)
        self._setDefault(probabilityCol="probability", rawPredictionCol="rawPrediction")

    def _transform(self, dataset):
        import torch
        fr = _Frame(dataset)
        m = self._margins(fr)
        b = self._xgb_sklearn_model
        if b.n_groups == 1:
            p1 = torch.sigmoid(m[:, 0])
            raw = torch.stack([-m[:, 0], m[:, 0]], 1)
            prob = torch.stack([1 - p1, p1], 1)
        else:
            raw, prob = m, torch.softmax(m, 1)
        gv = self.getOrDefault
        return fr.with_columns({
            gv("rawPredictionCol"): raw.numpy().astype(np.float64),
            gv("predictionCol"): prob.argmax(1).numpy().astype(np.float64),
            gv("probabilityCol"): prob.numpy().astype(np.float64)})


class XgboostRegressor(_XgboostEstimator):
    """
    XgboostRegressor is a Spark-ML style estimator implementing gradient-boosted regression trees on
    the B200-native GBDT engine.  It supports most parameters of ``xgboost.XGBRegressor`` (constructor,
    fit and predict); see the ``Param(parent=...`` entries.  Not supported: ``gpu_id`` (use ``use_gpu``),
    ``base_margin`` (use ``baseMarginCol``), ``validate_features``, ``output_margin``, ``eval_set`` (use
    ``validationIndicatorCol``), ``sample_weight[_eval_set]`` (use ``weightCol``).

    >>> from sparkdl.xgboost import XgboostRegressor, Vectors
    >>> import pandas as pd
    >>> df_train = pd.DataFrame({"features": [Vectors.dense(1.0, 2.0, 3.0), Vectors.sparse(3, {1: 1.0, 2: 5.5}),
    ...     Vectors.dense(4.0, 5.0, 6.0), Vectors.sparse(3, {1: 6.0, 2: 7.5})], "label": [0, 1, 2, 3],
    ...     "isVal": [False, False, True, True], "weight": [1.0, 2.0, 1.0, 2.0]})
    >>> reg = XgboostRegressor(max_depth=5, missing=0.0, validationIndicatorCol='isVal', weightCol='weight',
    ...     early_stopping_rounds=1, eval_metric='rmse')
    >>> model = reg.fit(df_train)
    >>> out = model.transform(df_train[["features"]])
    """

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def _model_class(self):
        return XgboostRegressorModel


class XgboostClassifier(_XgboostEstimator, HasProbabilityCol, HasRawPredictionCol):
    """
    XgboostClassifier is a Spark-ML style estimator implementing gradient-boosted classification trees
    (binary logistic / multi-class soft-prob).  ``rawPredictionCol`` always carries the predicted margins
    and ``probabilityCol`` the class probabilities.  Same unsupported arguments as :class:`XgboostRegressor`.

    >>> from sparkdl.xgboost import XgboostClassifier, Vectors
    >>> import pandas as pd
    >>> df_train = pd.DataFrame({"features": [Vectors.dense(1.0, 2.0, 3.0), Vectors.sparse(3, {1: 1.0, 2: 5.5}),
    ...     Vectors.dense(4.0, 5.0, 6.0), Vectors.sparse(3, {1: 6.0, 2: 7.5})], "label": [0, 1, 0, 1],
    ...     "isVal": [False, False, True, True], "weight": [1.0, 2.0, 1.0, 2.0]})
    >>> clf = XgboostClassifier(max_depth=5, missing=0.0, validationIndicatorCol='isVal', weightCol='weight',
    ...     early_stopping_rounds=1, eval_metric='logloss')
    >>> model = clf.fit(df_train)
    >>> out = model.transform(df_train[["features"]])
    """
    _is_classifier = True

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._setDefault(probabilityCol="probability", rawPredictionCol="rawPrediction")

    def _model_class(self):
        return XgboostClassifierModel


# ---- helpers ------------------------------------------------------------------------------------------------
def _spill_to_disk(X, digits):
    """``use_external_storage``: round to ``digits`` significant digits and keep the matrix in a
    memory-mapped file instead of RAM."""
    with np.errstate(divide="ignore", invalid="ignore"):
        mag = np.where(X != 0, np.floor(np.log10(np.abs(X))), 0)
    scale = np.power(10.0, digits - 1 - mag)
    Xr = np.where(np.isfinite(X), np.round(X * scale) / scale, X).astype(np.float32)
    fd, path = tempfile.mkstemp(prefix="sparkdl-xgb-", suffix=".bin")
    os.close(fd)
    mm = np.memmap(path, dtype=np.float32, mode="w+", shape=Xr.shape)
    mm[:] = Xr
    mm.flush()
    ro = np.memmap(path, dtype=np.float32, mode="r", shape=Xr.shape)
    del mm
    try:
        os.unlink(path)          # the mapping keeps the pages; no file is left behind when training ends (or crashes)
    except OSError:
        pass
    return ro


def _worker_train(shards, train_args):
    import sparkdl.hvd as hvd
    from spark_deep_learning_b200.xgboost import gbdt as _g
    hvd.init()
    r = hvd.rank()
    args = dict(train_args)
    args.update(shards[r])
    booster, evals = _g.train(**args)
    out = (booster.to_json(), evals) if r == 0 else None
    hvd.shutdown()
    return out


def _fit_distributed(train_args, num_workers, force_repartition, seed):
    """One worker per task slot through HorovodRunner; rows sharded, histograms allreduced
    (reference: 'Each XGBoost worker corresponds to one spark task', xgboost.py:58-64)."""
    from sparkdl import HorovodRunner
    X, y, w = train_args["X"], train_args["y"], train_args["weight"]
    n = len(y)
    idx = np.random.RandomState(seed).permutation(n) if force_repartition else np.arange(n)
    parts = np.array_split(idx, num_workers)
    ev = train_args["eval_set"]
    shards = []
    for r, rows in enumerate(parts):
        sh = {"X": np.ascontiguousarray(X[rows]), "y": y[rows], "weight": None if w is None else w[rows]}
        if ev is not None:
            vr = np.array_split(np.arange(len(ev[1])), num_workers)[r]
            sh["eval_set"] = (np.ascontiguousarray(ev[0][vr]), ev[1][vr], None if ev[2] is None else ev[2][vr])
        shards.append(sh)
    base = {k: v for k, v in train_args.items() if k not in ("X", "y", "weight", "eval_set")}
    js, evals = HorovodRunner(np=-num_workers).run(_worker_train, shards=shards, train_args=base)
    return gbdt.Booster.from_json(js), evals
