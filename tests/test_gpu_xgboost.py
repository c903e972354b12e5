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


@pytest.mark.parametrize("n,f,nodes", [(50000, 12, 1), (200001, 64, 32), (70000, 7, 48), (30000, 20, 64)])
def test_hist_kernels_bit_exact(n, f, nodes, monkeypatch):
    """shared-memory privatised histogram == global-atomic histogram == an index_add reference, bit for bit
    (nodes=64 exceeds the shared-memory tiling and must fall back silently)."""
    import torch
    from spark_deep_learning_b200.xgboost import gbdt
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(n)
    bins = torch.randint(0, 256, (n, f), device=dev, dtype=torch.uint8, generator=g)
    gh = torch.randint(-2 ** 30, 2 ** 30, (n, 2), device=dev, dtype=torch.int32, generator=g)
    pos = torch.randint(-1, nodes, (n,), device=dev, dtype=torch.int32, generator=g)
    ref = torch.zeros(nodes * f * 256, 2, dtype=torch.int64, device=dev)
    live = pos >= 0
    idx = (pos.long()[:, None] * f + torch.arange(f, device=dev)[None, :]) * 256 + bins.long()
    ref.index_add_(0, idx[live].reshape(-1), gh.long()[live][:, None, :].expand(-1, f, -1).reshape(-1, 2))
    out = {}
    for smem in ("1", "0"):
        monkeypatch.setenv("SPARKDL_GBDT_SMEM_HIST", smem)
        b = gbdt._HistBuilder(bins, gbdt._Dist(), use_fused=False)
        assert (b.binsT is not None) == (smem == "1")
        out[smem] = b.build(gh, pos, nodes).reshape(-1, 2).clone()
    assert torch.equal(out["1"], ref) and torch.equal(out["0"], ref)


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
