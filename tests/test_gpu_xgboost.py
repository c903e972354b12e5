"""GBDT kernels on the GPU: the int64 histogram kernel must reproduce the CPU path bit for bit, and the
fused hist-build + peer allreduce kernel (np=2) must give the same model as single-GPU training."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def _toy(n=20000, f=12, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.randn(n, f).astype(np.float32)
    y = (2 * X[:, 0] + np.sin(3 * X[:, 1]) + 1.5 * (X[:, 2] > 0.3) + 0.1 * rng.randn(n)).astype(np.float32)
    X[rng.rand(n, f) < 0.05] = np.nan
    return X, y


def test_gpu_training_matches_cpu():
    from sparkdl.xgboost import XgboostRegressor
    X, y = _toy()
    df = pd.DataFrame({"features": list(X), "label": y})
    kw = dict(n_estimators=10, max_depth=5, learning_rate=0.3)
    cpu = XgboostRegressor(**kw).fit(df)
    gpu = XgboostRegressor(use_gpu=True, **kw).fit(df)
    pc = cpu.transform(df)["prediction"].to_numpy()
    pg = gpu.transform(df)["prediction"].to_numpy()
    assert gpu.get_booster().attributes["hist_launches"] > 0
    assert np.allclose(pc, pg, atol=2e-3), np.abs(pc - pg).max()
    assert np.sqrt(np.mean((pg - y) ** 2)) < 0.6 * y.std()


@pytest.mark.multigpu
def test_distributed_gpu_fused_hist_allreduce():
    from sparkdl.xgboost import XgboostRegressor
    X, y = _toy(30000)
    df = pd.DataFrame({"features": list(X), "label": y})
    kw = dict(n_estimators=8, max_depth=5, learning_rate=0.3, use_gpu=True)
    one = XgboostRegressor(**kw).fit(df).transform(df)["prediction"].to_numpy()
    model = XgboostRegressor(num_workers=2, **kw).fit(df)
    two = model.transform(df)["prediction"].to_numpy()
    assert model.get_booster().attributes["workers"] == 2
    assert np.allclose(one, two, atol=2e-3), np.abs(one - two).max()
