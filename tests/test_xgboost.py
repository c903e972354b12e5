"""sparkdl.xgboost: the reference's doctests (xgboost.py:223-239, 311-326) as executable tests, the
param surface of SURVEY.md §2b, persistence, and data-parallel training on CPU workers."""
import numpy as np
import pandas as pd
import pytest

from sparkdl.xgboost import (Vectors, XgboostClassifier, XgboostClassifierModel, XgboostRegressor,
                             XgboostRegressorModel)


def _toy(n=3000, f=8, seed=0, nan_frac=0.1):
    rng = np.random.RandomState(seed)
    X = rng.randn(n, f).astype(np.float32)
    y = (2 * X[:, 0] + np.sin(3 * X[:, 1]) + 1.5 * (X[:, 2] > 0.3) + 0.1 * rng.randn(n)).astype(np.float32)
    X[rng.rand(n, f) < nan_frac] = np.nan
    return X, y


def test_reference_doctest_regressor():
    df_train = pd.DataFrame({
        "features": [Vectors.dense(1.0, 2.0, 3.0), Vectors.sparse(3, {1: 1.0, 2: 5.5}),
                     Vectors.dense(4.0, 5.0, 6.0), Vectors.sparse(3, {1: 6.0, 2: 7.5})],
        "label": [0, 1, 2, 3], "isVal": [False, False, True, True], "weight": [1.0, 2.0, 1.0, 2.0]})
    df_test = pd.DataFrame({"features": [Vectors.dense(1.0, 2.0, 3.0), Vectors.sparse(3, {1: 1.0, 2: 5.5})]})
    reg = XgboostRegressor(max_depth=5, missing=0.0, validationIndicatorCol='isVal', weightCol='weight',
                           early_stopping_rounds=1, eval_metric='rmse')
    model = reg.fit(df_train)
    assert isinstance(model, XgboostRegressorModel)
    out = model.transform(df_test)
    assert list(out.columns) == ["features", "prediction"] and len(out) == 2


def test_reference_doctest_classifier():
    df_train = pd.DataFrame({
        "features": [Vectors.dense(1.0, 2.0, 3.0), Vectors.sparse(3, {1: 1.0, 2: 5.5}),
                     Vectors.dense(4.0, 5.0, 6.0), Vectors.sparse(3, {1: 6.0, 2: 7.5})],
        "label": [0, 1, 0, 1], "isVal": [False, False, True, True], "weight": [1.0, 2.0, 1.0, 2.0]})
    clf = XgboostClassifier(max_depth=5, missing=0.0, validationIndicatorCol='isVal', weightCol='weight',
                            early_stopping_rounds=1, eval_metric='logloss')
    model = clf.fit(df_train)
    assert isinstance(model, XgboostClassifierModel)
    out = model.transform(pd.DataFrame({"features": [Vectors.dense(1.0, 2.0, 3.0)]}))
    assert {"rawPrediction", "prediction", "probability"} <= set(out.columns)
    assert abs(sum(out["probability"][0]) - 1.0) < 1e-6 and len(out["rawPrediction"][0]) == 2


def test_param_surface():
    reg = XgboostRegressor()
    for name in ("missing", "callbacks", "num_workers", "use_gpu", "force_repartition", "use_external_storage",
                 "external_storage_precision", "baseMarginCol", "featuresCol", "labelCol", "weightCol",
                 "predictionCol", "validationIndicatorCol", "max_depth", "n_estimators", "learning_rate"):
        assert reg.hasParam(name), name
    clf = XgboostClassifier()
    assert clf.hasParam("probabilityCol") and clf.hasParam("rawPredictionCol")
    assert np.isnan(reg.getOrDefault("missing")) and reg.getOrDefault("use_external_storage") is False
    with pytest.raises(TypeError):
        XgboostRegressor(num_workers="two")           # toInt converter (ref xgboost.py:63)
    for bad in ("gpu_id", "base_margin", "validate_features", "output_margin", "eval_set", "sample_weight",
                "sample_weight_eval_set"):
        with pytest.raises(ValueError, match="Unsupported param"):
            XgboostRegressor(**{bad: 1})
    with pytest.raises(ValueError, match="Invalid param name"):
        XgboostRegressor(no_such_param=1)
    c = reg.copy({reg.max_depth: 3})
    assert c.getOrDefault("max_depth") == 3 and reg.getOrDefault("max_depth") == 6
    assert "max_depth" in reg.explainParams()


def test_regression_quality_missing_weights_early_stopping():
    X, y = _toy()
    rng = np.random.RandomState(1)
    df = pd.DataFrame({"features": list(X), "label": y, "isVal": rng.rand(len(y)) < 0.2,
                       "w": rng.uniform(0.5, 2.0, len(y))})
    m = XgboostRegressor(n_estimators=80, max_depth=5, learning_rate=0.2, validationIndicatorCol="isVal",
                         weightCol="w", early_stopping_rounds=5, subsample=0.9, colsample_bytree=0.9).fit(df)
    hist = m.evals_result()["validation_0"]["rmse"]
    assert hist[-1] < 0.45 * hist[0]
    b = m.get_booster()
    assert b.best_iteration is not None and b.num_boosted_rounds <= 80
    pred = m.transform(df)["prediction"].to_numpy()
    assert np.sqrt(np.mean((pred - y) ** 2)) < 0.5 * y.std()
    assert len(b.get_dump()) == len(b.trees) and b.get_score()


def test_binary_and_multiclass():
    X, y = _toy(2000)
    yb = (y > np.median(y)).astype(np.float32)
    out = XgboostClassifier(n_estimators=25, max_depth=4).fit(pd.DataFrame({"features": list(X), "label": yb})) \
        .transform(pd.DataFrame({"features": list(X)}))
    assert (out["prediction"].to_numpy() == yb).mean() > 0.95
    raw = np.stack(out["rawPrediction"].to_list())
    assert np.allclose(raw[:, 0], -raw[:, 1])
    y3 = np.digitize(y, np.quantile(y, [0.33, 0.66])).astype(np.float32)
    m3 = XgboostClassifier(n_estimators=20, max_depth=4).fit(pd.DataFrame({"features": list(X), "label": y3}))
    o3 = m3.transform(pd.DataFrame({"features": list(X)}))
    assert (o3["prediction"].to_numpy() == y3).mean() > 0.85 and len(o3["probability"][0]) == 3


def test_persistence_warm_start_callbacks_base_margin(tmp_path):
    X, y = _toy(1500)
    df = pd.DataFrame({"features": list(X), "label": y, "bm": np.full(len(y), 0.25, dtype=np.float32)})
    seen = []
    est = XgboostRegressor(n_estimators=10, max_depth=3, callbacks=[lambda env: seen.append(env["iteration"]) or False],
                           baseMarginCol="bm")
    model = est.fit(df)
    assert seen == list(range(10))
    p = model.transform(df)["prediction"].to_numpy()
    model.write().overwrite().save(str(tmp_path / "model"))
    again = XgboostRegressorModel.load(str(tmp_path / "model"))
    assert np.allclose(again.transform(df)["prediction"].to_numpy(), p)
    est.write().overwrite().save(str(tmp_path / "est"))
    est2 = XgboostRegressor.load(str(tmp_path / "est"))
    assert est2.getOrDefault("max_depth") == 3 and callable(est2.getOrDefault("callbacks")[0])
    with pytest.raises(IOError):
        model.write().save(str(tmp_path / "model"))
    # warm start from get_booster() (ref xgboost.py:198-199)
    more = XgboostRegressor(n_estimators=5, max_depth=3, xgb_model=model.get_booster()).fit(df[["features", "label"]])
    assert more.get_booster().num_boosted_rounds == 15


def test_external_storage_rounding():
    X, y = _toy(800, nan_frac=0.0)
    df = pd.DataFrame({"features": list(X), "label": y})
    m = XgboostRegressor(n_estimators=10, use_external_storage=True, external_storage_precision=3).fit(df)
    assert np.isfinite(m.transform(df)["prediction"].to_numpy()).all()
    with pytest.raises(ValueError):
        XgboostRegressor(use_external_storage=True, weightCol="label").fit(df)


def test_distributed_cpu_workers_match_single_worker(monkeypatch):
    """num_workers=2: rows sharded over two HorovodRunner workers, histograms allreduced as exact
    int64 sums -> the same model as single-worker training."""
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")
    X, y = _toy(1200, f=6)
    df = pd.DataFrame({"features": list(X), "label": y})
    kw = dict(n_estimators=8, max_depth=4, learning_rate=0.3)
    single = XgboostRegressor(**kw).fit(df).transform(df)["prediction"].to_numpy()
    dist_model = XgboostRegressor(num_workers=2, force_repartition=True, **kw).fit(df)
    dist = dist_model.transform(df)["prediction"].to_numpy()
    assert dist_model.get_booster().attributes["workers"] == 2
    assert np.allclose(single, dist, atol=1e-4), np.abs(single - dist).max()
    with pytest.raises(ValueError, match="distributed"):
        XgboostRegressor(num_workers=2, baseMarginCol="label").fit(df)


def test_pipeline_and_cross_validator():
    """ref xgboost.py:168-169: usable in Pipeline and CrossValidator (copy / param maps)."""
    from spark_deep_learning_b200.xgboost.ml import (CrossValidator, ParamGridBuilder, Pipeline,
                                                     RegressionEvaluator)
    X, y = _toy(900, f=5, nan_frac=0.0)
    df = pd.DataFrame({"features": list(X), "label": y})
    reg = XgboostRegressor(n_estimators=15, learning_rate=0.3)
    grid = ParamGridBuilder().addGrid(reg.max_depth, [1, 4]).build()
    cv = CrossValidator(reg, grid, RegressionEvaluator(), numFolds=2).fit(df)
    assert cv.bestIndex == 1 and cv.avgMetrics[1] < cv.avgMetrics[0]
    out = Pipeline([reg]).fit(df).transform(df)
    assert "prediction" in out.columns


def test_scale_pos_weight_and_ignored_param_warning():
    X, y = _toy(1500)
    yb = (y > np.quantile(y, 0.9)).astype(np.float32)          # 10 % positives
    df = pd.DataFrame({"features": list(X), "label": yb})
    plain = XgboostClassifier(n_estimators=15, max_depth=3).fit(df).transform(df)
    heavy = XgboostClassifier(n_estimators=15, max_depth=3, scale_pos_weight=8.0).fit(df).transform(df)
    rec = lambda o: ((o["prediction"].to_numpy() == 1) & (yb == 1)).sum() / max(1, (yb == 1).sum())  # noqa: E731
    assert rec(heavy) >= rec(plain)
    with pytest.warns(UserWarning, match="not used by the B200 GBDT engine"):
        XgboostRegressor(n_estimators=2, booster="dart").fit(pd.DataFrame({"features": list(X), "label": y}))


def _three_class_frame(n=900, seed=3):
    import pandas as pd
    rng = np.random.RandomState(seed)
    X = rng.randn(n, 6).astype(np.float32)
    y = np.argmax(X[:, :3] + 0.3 * rng.randn(n, 3), axis=1).astype(np.float32)
    return pd.DataFrame({"features": list(X), "label": y})


def test_train_validation_split_picks_the_better_map():
    from spark_deep_learning_b200.xgboost.ml import MulticlassClassificationEvaluator, ParamGridBuilder, TrainValidationSplit
    df = _three_class_frame()
    est = XgboostClassifier(n_estimators=8, learning_rate=0.3)
    grid = ParamGridBuilder().addGrid(est.getParam("max_depth"), [1, 4]).build()
    m = TrainValidationSplit(est, grid, MulticlassClassificationEvaluator(), trainRatio=0.7, seed=1).fit(df)
    assert len(m.validationMetrics) == 2 and m.bestIndex == int(np.argmax(m.validationMetrics))
    acc = float((m.transform(df)["prediction"] == df["label"]).mean())
    assert acc > 0.8


def test_one_vs_rest_matches_native_multiclass_quality():
    from spark_deep_learning_b200.xgboost.ml import OneVsRest
    df = _three_class_frame()
    ovr = OneVsRest(XgboostClassifier(n_estimators=12, max_depth=3, learning_rate=0.3)).fit(df)
    out = ovr.transform(df)
    assert len(ovr.models) == 3 and np.stack(out["rawPrediction"].to_list()).shape == (len(df), 3)
    acc_ovr = float((out["prediction"] == df["label"]).mean())
    native = XgboostClassifier(n_estimators=12, max_depth=3, learning_rate=0.3).fit(df).transform(df)
    acc_native = float((native["prediction"] == df["label"]).mean())
    assert acc_ovr > 0.85 and abs(acc_ovr - acc_native) < 0.08


def test_binary_evaluator_auc():
    import pandas as pd
    from spark_deep_learning_b200.xgboost.ml import BinaryClassificationEvaluator
    ev = BinaryClassificationEvaluator()
    df = pd.DataFrame({"rawPrediction": [[-2.0, 2.0], [-1.0, 1.0], [1.0, -1.0], [0.5, -0.5]], "label": [1, 1, 0, 0]})
    assert ev.evaluate(df) == 1.0
    df2 = pd.DataFrame({"rawPrediction": [[0, 1.0], [0, 1.0], [0, 1.0], [0, 1.0]], "label": [1, 0, 1, 0]})
    assert ev.evaluate(df2) == 0.5                              # all ties
    rng = np.random.RandomState(0)
    s, y = rng.randn(300), rng.rand(300) < 0.4
    s[y] += 1.0
    pairs = (s[y][:, None] > s[~y][None, :]).mean()
    assert abs(ev.evaluate(pd.DataFrame({"rawPrediction": [[-v, v] for v in s], "label": y.astype(float)})) - pairs) < 1e-9


def test_sibling_subtraction_and_leaf_skip_are_bit_exact(monkeypatch, tmp_path):
    """Building only the lighter child (+ parent - child) and taking leaf totals from the parent's split must give
    the very same model as building every node of every level (int64 fixed-point histograms are exact)."""
    from spark_deep_learning_b200.xgboost import gbdt
    rng = np.random.RandomState(5)
    X = rng.randn(3000, 9).astype(np.float32)
    X[rng.rand(*X.shape) < 0.1] = np.nan
    y = (np.nan_to_num(X[:, 0]) * np.nan_to_num(X[:, 1]) + 0.2 * rng.randn(3000)).astype(np.float32)
    dumps, launches = {}, {}
    for sub in ("0", "1"):
        for skip in ("0", "1"):
            monkeypatch.setenv("SPARKDL_GBDT_SIBLING_SUBTRACT", sub)
            monkeypatch.setenv("SPARKDL_GBDT_SKIP_LEAF_HIST", skip)
            b, _ = gbdt.train(dict(n_estimators=5, max_depth=4, subsample=0.8, colsample_bytree=0.7), X, y)
            path = str(tmp_path / ("m%s%s.json" % (sub, skip)))
            b.save_model(path)
            dumps[(sub, skip)] = open(path).read()
    assert len(set(dumps.values())) == 1


def test_native_and_torch_cpu_histograms_give_the_same_model(monkeypatch, tmp_path):
    from spark_deep_learning_b200.xgboost import gbdt
    rng = np.random.RandomState(11)
    X = rng.randn(4000, 7).astype(np.float32)
    y = (X[:, 0] - X[:, 3] * X[:, 4] + 0.1 * rng.randn(4000)).astype(np.float32)
    dumps = []
    for native in ("1", "0"):
        monkeypatch.setenv("SPARKDL_GBDT_NATIVE_CPU", native)
        monkeypatch.setitem(gbdt._CORE_HIST, "fn", False)           # re-resolve under the new setting
        b, _ = gbdt.train(dict(n_estimators=4, max_depth=5), X, y)
        assert (gbdt._core_hist_fn() is not None) == (native == "1")
        path = str(tmp_path / ("m%s.json" % native))
        b.save_model(path)
        dumps.append(open(path).read())
    monkeypatch.setitem(gbdt._CORE_HIST, "fn", False)
    assert dumps[0] == dumps[1]


def _wiggly(n=3000, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, 4).astype(np.float32)
    y = (5 * X[:, 0] + np.sin(10 * np.pi * X[:, 0]) - 5 * X[:, 1] - np.cos(10 * np.pi * X[:, 1]) + 0.3 * rng.randn(n)).astype(np.float32)
    return X, y


@pytest.mark.parametrize("spec", ["(1,-1)", [1, -1, 0, 0], {0: 1, 1: -1}])
def test_monotone_constraints_are_enforced(spec):
    """xgboost's monotone_constraints: predictions never decrease in feature 0 / never increase in feature 1."""
    X, y = _wiggly()
    df = pd.DataFrame({"features": list(X), "label": y})
    model = XgboostRegressor(n_estimators=25, max_depth=5, monotone_constraints=spec).fit(df)
    free = XgboostRegressor(n_estimators=25, max_depth=5).fit(df)
    grid = np.linspace(0, 1, 150, dtype=np.float32)
    rng = np.random.RandomState(1)

    def violations(m, f, sign):
        bad = 0
        for _ in range(8):
            Z = np.tile(rng.rand(4).astype(np.float32), (len(grid), 1))
            Z[:, f] = grid
            pr = m.transform(pd.DataFrame({"features": list(Z)}))["prediction"].to_numpy()
            bad += int((np.diff(pr) * sign < -1e-6).sum())
        return bad
    assert violations(model, 0, 1) == 0 and violations(model, 1, -1) == 0
    assert violations(free, 0, 1) > 0                       # the unconstrained fit follows the sine
    with pytest.raises(ValueError):
        XgboostRegressor(monotone_constraints="(2,0)").fit(df)


def test_max_delta_step_and_column_sampling_below_the_tree_level():
    from spark_deep_learning_b200.xgboost import gbdt
    X, y = _wiggly(2000, seed=2)
    b, _ = gbdt.train(dict(n_estimators=3, max_depth=4, learning_rate=1.0, max_delta_step=0.25), X, y)
    leaves = [abs(v) for t in b.trees for v, l in zip(t.value, t.left) if l < 0]
    assert max(leaves) <= 0.25 + 1e-9 and max(leaves) > 0.2
    base, _ = gbdt.train(dict(n_estimators=6, max_depth=4), X, y)
    for kw in ({"colsample_bylevel": 0.5}, {"colsample_bynode": 0.5}):
        m1, _ = gbdt.train(dict(n_estimators=6, max_depth=4, random_state=3, **kw), X, y)
        m2, _ = gbdt.train(dict(n_estimators=6, max_depth=4, random_state=3, **kw), X, y)
        assert [t.feature for t in m1.trees] == [t.feature for t in m2.trees]            # seeded -> reproducible
        assert [t.feature for t in m1.trees] != [t.feature for t in base.trees]          # and it does something
        pr = np.asarray(m1.predict(X)).ravel()
        assert np.sqrt(np.mean((pr - y) ** 2)) < 1.5


def test_distributed_workers_agree_with_constraints_and_node_sampling(monkeypatch):
    """The per-level / per-node column draws and the monotone bounds are functions of allreduced data and a shared
    seed, so two workers still produce the single-worker model."""
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")
    X, y = _wiggly(1200, seed=4)
    df = pd.DataFrame({"features": list(X), "label": y})
    kw = dict(n_estimators=6, max_depth=4, monotone_constraints="(1,-1)", colsample_bynode=0.75, colsample_bylevel=0.75,
              max_delta_step=2.0)
    single = XgboostRegressor(**kw).fit(df).transform(df)["prediction"].to_numpy()
    dist = XgboostRegressor(num_workers=2, force_repartition=True, **kw).fit(df).transform(df)["prediction"].to_numpy()
    assert np.allclose(single, dist, atol=1e-4), np.abs(single - dist).max()


@pytest.mark.parametrize("name", ["reg:squaredlogerror", "reg:pseudohubererror", "reg:gamma", "reg:tweedie", "binary:logitraw"])
def test_objective_gradients_match_autograd(name):
    """grad / hess of every closed-form objective equal d/dm and d2/dm2 of its loss (fp64 autograd as the reference)."""
    import torch
    from spark_deep_learning_b200.xgboost import gbdt
    rho = 1.5
    losses = {
        "reg:squaredlogerror": lambda m, y: 0.5 * (torch.log1p(m) - torch.log1p(y)) ** 2,
        "reg:pseudohubererror": lambda m, y: torch.sqrt(1 + (m - y) ** 2) - 1,
        "reg:gamma": lambda m, y: y / torch.exp(m) + m,
        "reg:tweedie": lambda m, y: -y * torch.exp((1 - rho) * m) / (1 - rho) + torch.exp((2 - rho) * m) / (2 - rho),
        "binary:logitraw": lambda m, y: torch.nn.functional.softplus(m) - y * m,
    }
    g0 = torch.Generator().manual_seed(0)
    m = (torch.rand(40, dtype=torch.float64, generator=g0) * 2 + 0.1).requires_grad_()
    y = (torch.rand(40, dtype=torch.float64, generator=g0) > 0.5).double() if name.startswith("binary") \
        else torch.rand(40, dtype=torch.float64, generator=g0) * 3 + 0.1
    g, = torch.autograd.grad(losses[name](m, y).sum(), m, create_graph=True)
    h = torch.stack([torch.autograd.grad(g[i], m, retain_graph=True)[0][i] for i in range(40)])
    G, H = gbdt._Objective(name, 1).grad_hess(m.detach().unsqueeze(1), y, torch.ones(40, dtype=torch.float64))
    assert torch.allclose(G[:, 0], g.detach(), atol=1e-12) and torch.allclose(H[:, 0], h, atol=1e-12)


def test_poisson_regression_through_the_estimator():
    rng = np.random.RandomState(0)
    X = rng.rand(3000, 4).astype(np.float32)
    mu = np.exp(0.5 + 1.5 * X[:, 0] - X[:, 1])
    df = pd.DataFrame({"features": list(X), "label": rng.poisson(mu).astype(np.float32), "isVal": rng.rand(3000) < 0.2})
    model = XgboostRegressor(objective="count:poisson", n_estimators=40, max_depth=3, learning_rate=0.2,
                             validationIndicatorCol="isVal").fit(df)
    pred = model.transform(df)["prediction"].to_numpy()
    assert pred.min() > 0 and np.corrcoef(pred, mu)[0, 1] > 0.97 and abs(pred.mean() / mu.mean() - 1) < 0.1
    hist = model._evals_result["validation_0"]["poisson-nloglik"]
    assert hist[-1] < hist[0]
    with pytest.raises(Exception, match="non-negative"):
        XgboostRegressor(objective="count:poisson", n_estimators=1).fit(pd.DataFrame({"features": list(X), "label": -mu}))


def test_interaction_constraints_limit_the_features_on_every_path():
    """xgboost's interaction_constraints: features on one root-to-leaf path all come from one group."""
    from spark_deep_learning_b200.xgboost import gbdt
    rng = np.random.RandomState(0)
    X = rng.rand(3000, 5).astype(np.float32)
    y = (X[:, 0] * X[:, 2] + X[:, 1] * X[:, 3] + X[:, 4] + 0.05 * rng.randn(3000)).astype(np.float32)

    def paths(t):
        out = []

        def rec(i, acc):
            if t.left[i] < 0:
                out.append(frozenset(acc))
                return
            rec(t.left[i], acc | {t.feature[i]})
            rec(t.right[i], acc | {t.feature[i]})
        rec(0, frozenset())
        return out
    free, _ = gbdt.train(dict(n_estimators=10, max_depth=4), X, y)
    assert any(len(p) > 2 for t in free.trees for p in paths(t))
    for spec in ("[[0, 2], [1, 3, 4]]", [[0, 1], [2, 3]]):
        groups = gbdt.parse_interactions(spec, 5)
        b, _ = gbdt.train(dict(n_estimators=10, max_depth=4, interaction_constraints=spec), X, y)
        for t in b.trees:
            for p in paths(t):
                assert len(p) <= 1 or any(p <= g for g in groups), sorted(p)
    with pytest.raises(ValueError):
        gbdt.train(dict(n_estimators=1, interaction_constraints=[[0, 9]]), X, y)


def test_distributed_auc_early_stopping_is_consistent(monkeypatch):
    """AUC is a rank statistic: the workers pool their validation rows so that all of them see the same value and
    take the same early-stopping decision (a per-shard AUC would let them diverge and hang)."""
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")
    rng = np.random.RandomState(7)
    X = rng.randn(1500, 6).astype(np.float32)
    yb = ((X[:, 0] + 0.5 * X[:, 1] + 0.8 * rng.randn(1500)) > 0).astype(np.float32)
    df = pd.DataFrame({"features": list(X), "label": yb, "isVal": rng.rand(1500) < 0.3})
    kw = dict(n_estimators=60, max_depth=3, learning_rate=0.5, eval_metric="auc", early_stopping_rounds=3,
              validationIndicatorCol="isVal")
    single = XgboostClassifier(**kw).fit(df)
    dist = XgboostClassifier(num_workers=2, **kw).fit(df)
    bs, bd = single.get_booster(), dist.get_booster()
    assert bd.num_boosted_rounds < 60                                  # it did stop early
    assert bs.best_iteration == bd.best_iteration and abs(bs.best_score - bd.best_score) < 1e-6
    assert np.allclose(single._evals_result["validation_0"]["auc"], dist._evals_result["validation_0"]["auc"], atol=1e-6)


def test_num_parallel_tree_boosted_forest(tmp_path):
    """num_parallel_tree: several trees per round on their own row samples, leaf values scaled by 1/k; rounds,
    early-stopping limits and persistence count trees_per_round."""
    from spark_deep_learning_b200.xgboost import gbdt
    rng = np.random.RandomState(0)
    X = rng.rand(2500, 6).astype(np.float32)
    y = (np.sin(6 * X[:, 0]) + X[:, 1] * X[:, 2] + 0.3 * rng.randn(2500)).astype(np.float32)
    Xt = rng.rand(1500, 6).astype(np.float32)
    yt = np.sin(6 * Xt[:, 0]) + Xt[:, 1] * Xt[:, 2]
    one, _ = gbdt.train(dict(n_estimators=1, learning_rate=1.0, max_depth=6), X, y)
    forest, _ = gbdt.train(dict(n_estimators=1, learning_rate=1.0, max_depth=6, num_parallel_tree=20, subsample=0.6,
                                colsample_bynode=0.7), X, y)
    rmse = lambda b: float(np.sqrt(np.mean((np.asarray(b.predict(Xt)).ravel() - yt) ** 2)))  # noqa: E731
    assert len(forest.trees) == 20 and forest.num_boosted_rounds == 1 and rmse(forest) < rmse(one)
    df = pd.DataFrame({"features": list(X), "label": y, "isVal": rng.rand(2500) < 0.25})
    m = XgboostRegressor(n_estimators=30, num_parallel_tree=3, subsample=0.7, max_depth=3, early_stopping_rounds=2,
                         validationIndicatorCol="isVal", learning_rate=0.8).fit(df)
    b = m.get_booster()
    assert len(b.trees) == 3 * b.num_boosted_rounds and b.trees_per_round == 3
    m.write().overwrite().save(str(tmp_path / "rf"))
    again = XgboostRegressorModel.load(str(tmp_path / "rf"))
    assert again.get_booster().trees_per_round == 3
    assert np.allclose(again.transform(df)["prediction"], m.transform(df)["prediction"])


def test_gbdt_fuzz_three_workers_equal_single_worker(monkeypatch):
    """Ten seeded random configurations (objective, depth, regularisation, column sampling, coarse max_bin, monotone /
    interaction constraints, weights, NaN or 0.0 as missing): 3 row-sharded workers build the single-worker trees."""
    import random
    from sparkdl import HorovodRunner
    from spark_deep_learning_b200.xgboost import gbdt
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")
    rnd = random.Random(5)
    cfgs = []
    for _ in range(10):
        obj = rnd.choice(["reg:squarederror", "binary:logistic", "multi:softprob", "count:poisson", "reg:pseudohubererror"])
        p = dict(objective=obj, n_estimators=rnd.randint(2, 4), max_depth=rnd.randint(1, 5), learning_rate=rnd.choice([0.1, 0.3, 1.0]),
                 reg_lambda=rnd.choice([0.0, 1.0, 5.0]), reg_alpha=rnd.choice([0.0, 0.5]), gamma=rnd.choice([0.0, 0.1]),
                 min_child_weight=rnd.choice([0.0, 1.0, 5.0]), colsample_bytree=rnd.choice([1.0, 0.6]),
                 colsample_bynode=rnd.choice([1.0, 0.7]), max_bin=rnd.choice([256, 16, 4]), max_delta_step=rnd.choice([0.0, 0.5]),
                 random_state=rnd.randint(0, 99))
        if rnd.random() < 0.3:
            p["monotone_constraints"] = [rnd.choice([-1, 0, 1]) for _ in range(5)]
        if rnd.random() < 0.3:
            p["interaction_constraints"] = [[0, 1], [2, 3, 4]]
        if obj == "multi:softprob":
            p["num_class"] = 3
        cfgs.append((p, rnd.randint(0, 10 ** 6), rnd.choice([float("nan"), 0.0])))

    def make(p, dseed, missing):
        import numpy as np
        rng = np.random.RandomState(dseed)
        n = rng.randint(50, 500)
        X = rng.randn(n, 5).astype(np.float32)
        X[rng.rand(n, 5) < 0.15] = missing
        base = np.nan_to_num(X[:, 0]) - np.nan_to_num(X[:, 1]) * np.nan_to_num(X[:, 2])
        obj = p["objective"]
        if obj == "binary:logistic":
            y = (base > 0).astype(np.float32)
        elif obj == "multi:softprob":
            y = np.digitize(base, [-0.5, 0.5]).astype(np.float32)
        elif obj == "count:poisson":
            y = rng.poisson(np.exp(np.clip(base, -2, 2))).astype(np.float32)
        else:
            y = (base + 0.1 * rng.randn(n)).astype(np.float32)
        return X, y, rng.uniform(0.5, 2.0, n).astype(np.float32)

    def worker(cfgs):
        import numpy as np
        import sparkdl.hvd as hvd
        from spark_deep_learning_b200.xgboost import gbdt as g
        hvd.init()
        outs = []
        for p, dseed, missing in cfgs:
            X, y, w = make(p, dseed, missing)
            idx = np.array_split(np.arange(len(y)), hvd.size())[hvd.rank()]
            outs.append(g.train(p, X[idx], y[idx], weight=w[idx], missing=missing)[0].to_json())
        return outs
    dist = HorovodRunner(np=-3).run(worker, cfgs=cfgs)
    for (p, dseed, missing), jd in zip(cfgs, dist):
        X, y, w = make(p, dseed, missing)
        bs = gbdt.train(p, X, y, weight=w, missing=missing)[0]
        bd = gbdt.Booster.from_json(jd)
        assert [t.feature for t in bs.trees] == [t.feature for t in bd.trees], p
        ps, pd_ = np.asarray(bs.predict(X, output_margin=True)), np.asarray(bd.predict(X, output_margin=True))
        assert np.abs(ps - pd_).max() < 1e-4, p


def test_accuracy_is_on_par_with_sklearn_hist_gradient_boosting():
    """Independent reference (xgboost itself is not installable offline): same depth / rounds / learning rate / L2 as
    scikit-learn's histogram GBDT must give comparable held-out error on regression and classification."""
    from sklearn.ensemble import HistGradientBoostingClassifier, HistGradientBoostingRegressor
    from spark_deep_learning_b200.xgboost import gbdt
    f = lambda X: 2 * np.sin(X[:, 0]) + X[:, 1] * X[:, 2] + (X[:, 3] > 0.5) * 1.5 + 0.5 * X[:, 4] ** 2  # noqa: E731
    rng = np.random.RandomState(2)
    X, Xt = rng.randn(4000, 8).astype(np.float32), rng.randn(2000, 8).astype(np.float32)
    y = (f(X) + 0.3 * rng.randn(4000)).astype(np.float32)
    common = dict(max_leaf_nodes=None, min_samples_leaf=1, l2_regularization=1.0, early_stopping=False)
    sk = HistGradientBoostingRegressor(max_iter=60, learning_rate=0.1, max_depth=5, **common).fit(X, y)
    us, _ = gbdt.train(dict(n_estimators=60, learning_rate=0.1, max_depth=5), X, y)
    r_sk = np.sqrt(np.mean((sk.predict(Xt) - f(Xt)) ** 2))
    r_us = np.sqrt(np.mean((np.asarray(us.predict(Xt)).ravel() - f(Xt)) ** 2))
    assert r_us < 1.15 * r_sk, (r_us, r_sk)
    yb = (f(X) + 0.5 * rng.randn(4000) > 1.0).astype(np.float32)
    skc = HistGradientBoostingClassifier(max_iter=40, learning_rate=0.2, max_depth=4, **common).fit(X, yb)
    usc, _ = gbdt.train(dict(objective="binary:logistic", n_estimators=40, learning_rate=0.2, max_depth=4), X, yb)
    a_sk = (skc.predict(Xt) == (f(Xt) > 1.0)).mean()
    a_us = ((np.asarray(usc.predict(Xt)).ravel() > 0.5) == (f(Xt) > 1.0)).mean()
    assert a_us > a_sk - 0.02, (a_us, a_sk)


def test_spark_dataframe_adapter_with_a_duck_typed_frame():
    """pyspark is not installable here: exercise the Spark branch of the dataset adapter (toPandas in, createDataFrame
    out, vector objects with .toArray()) with a stand-in that has exactly that interface."""
    class Vec(object):
        def __init__(self, values):
            self._v = np.asarray(values, dtype=np.float64)

        def toArray(self):
            return self._v

    class Session(object):
        def createDataFrame(self, pdf):
            return FakeSparkFrame(pdf)

    class FakeSparkFrame(object):
        sparkSession = Session()

        def __init__(self, pdf):
            self._pdf = pdf

        def toPandas(self):
            return self._pdf.copy()
    X, y = _toy(400, f=5)
    sdf = FakeSparkFrame(pd.DataFrame({"features": [Vec(r) for r in X], "label": y}))
    model = XgboostRegressor(n_estimators=10, max_depth=3).fit(sdf)
    out = model.transform(sdf)
    assert isinstance(out, FakeSparkFrame)
    pdf = out.toPandas()
    ref = XgboostRegressor(n_estimators=10, max_depth=3).fit(pd.DataFrame({"features": list(X), "label": y}))
    assert np.allclose(pdf["prediction"], ref.transform(pd.DataFrame({"features": list(X)}))["prediction"])


def test_real_values_never_land_in_the_missing_bin():
    """with max_bin=256 the top quantile used to be binned as 255 == MISSING_BIN, so training
    routed those rows by default_left while inference routed them by value."""
    import torch
    from spark_deep_learning_b200.xgboost import gbdt
    g = torch.Generator().manual_seed(0)
    X = torch.rand(5000, 3, generator=g)
    dist = gbdt._Dist()
    cuts = gbdt.compute_cuts(X, float("nan"), 256, dist)
    assert max(c.numel() for c in cuts) <= gbdt.MISSING_BIN - 1
    bins = gbdt.bin_features(X, cuts, float("nan"))
    assert int((bins == gbdt.MISSING_BIN).sum()) == 0
    X[::7, 1] = float("nan")
    bins = gbdt.bin_features(X, cuts, float("nan"))
    assert int((bins[:, 1] == gbdt.MISSING_BIN).sum()) == int(torch.isnan(X[:, 1]).sum())
    assert int((bins[:, 0] == gbdt.MISSING_BIN).sum()) == 0


def test_get_score_importance_types_follow_xgboost_semantics():
    from spark_deep_learning_b200.xgboost import gbdt
    rng = np.random.RandomState(3)
    X = rng.randn(3000, 6).astype(np.float32)
    y = (3 * X[:, 0] + np.sin(2 * X[:, 1]) + 0.05 * rng.randn(3000)).astype(np.float32)     # features 2..5 are noise
    b, _ = gbdt.train(dict(n_estimators=10, max_depth=4, learning_rate=0.3), X, y)
    w, g, c = b.get_score(), b.get_score(importance_type="gain"), b.get_score(importance_type="cover")
    tg, tc = b.get_score(importance_type="total_gain"), b.get_score(importance_type="total_cover")
    assert set(w) == set(g) == set(c) == set(tg) == set(tc) and "f0" in w
    for k in w:
        assert w[k] >= 1 and abs(g[k] * w[k] - tg[k]) <= 1e-4 * max(1.0, tg[k]) and abs(c[k] * w[k] - tc[k]) <= 1e-4 * max(1.0, tc[k])
        assert 0 < c[k] <= 3000 + 1e-3                       # squared error: hessian mass = number of rows in the node
    assert max(tg, key=tg.get) == "f0"                        # the dominant signal carries the largest total gain
    from sparkdl.xgboost import XgboostRegressor
    m = XgboostRegressor(n_estimators=5, max_depth=3, importance_type="total_gain").fit(
        pd.DataFrame({"features": list(X[:500]), "label": y[:500]}))
    assert int(np.argmax(m.feature_importances_)) == 0 and m.get_feature_importances("weight")["f0"] >= 1
    imp = b.feature_importances("total_gain")
    assert imp.shape == (6,) and abs(float(imp.sum()) - 1.0) < 1e-5 and int(imp.argmax()) == 0
    with pytest.raises(ValueError):
        b.get_score(importance_type="nope")


def test_booster_round_trips_through_the_xgboost_json_schema(tmp_path):
    """ref xgboost.py:130-133: get_booster() is an xgboost Booster; interoperability is the JSON model schema."""
    import json
    import torch
    from spark_deep_learning_b200.xgboost import gbdt
    X, y = _toy(2000)
    df = pd.DataFrame({"features": list(X), "label": (y > y.mean()).astype(np.float32)})
    for est, kw in ((XgboostRegressor, {}), (XgboostClassifier, {})):
        model = est(n_estimators=6, max_depth=4, missing=0.0 if est is XgboostClassifier else float("nan"), **kw).fit(df)
        b = model.get_booster()
        path = str(tmp_path / ("m_%s.json" % est.__name__))
        b.save_model(path)
        doc = json.load(open(path))
        gbm = doc["learner"]["gradient_booster"]["model"]
        assert doc["version"][0] >= 1 and doc["learner"]["gradient_booster"]["name"] == "gbtree"
        assert int(gbm["gbtree_model_param"]["num_trees"]) == len(b.trees) == len(gbm["trees"]) == len(gbm["tree_info"])
        t0 = gbm["trees"][0]
        for key in ("left_children", "right_children", "parents", "split_indices", "split_conditions", "default_left",
                    "base_weights", "loss_changes", "sum_hessian", "split_type", "tree_param"):
            assert key in t0
        assert t0["parents"][0] == 2147483647 and int(t0["tree_param"]["num_nodes"]) == len(t0["left_children"])
        leaf = t0["left_children"].index(-1)
        assert t0["right_children"][leaf] == -1 and t0["split_conditions"][leaf] == t0["base_weights"][leaf]
        assert doc["learner"]["objective"]["name"] == b.objective
        b2 = gbdt.Booster.load_model(path)
        Xt = torch.from_numpy(X)
        assert torch.allclose(b.predict(Xt), b2.predict(Xt), atol=1e-6)
        assert b2.objective == b.objective and b2.n_features == b.n_features
        # the native format still loads
        p2 = str(tmp_path / "native.bin")
        b.save_model(p2)
        assert torch.allclose(gbdt.Booster.load_model(p2).predict(Xt), b.predict(Xt))
