#!/usr/bin/env python
"""BASELINE.json config 4: the sparkdl.xgboost allreduce path on synthetic tabular data under HorovodRunner(np=-N).

Each worker generates its own row shard on its GPU (rows_per_worker x features), trains `rounds` boosting rounds of
depth `depth`, and reports (a) seconds per boosting round for the full trainer and (b) microseconds per tree level of
the fused histogram-build + peer-memory allreduce kernel (K8) next to the unfused pair (histogram kernel, then the
generic int64 allreduce).  Device-timed with CUDA events, max over ranks."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(rows=1000000, features=64, rounds=5, depth=6):
    import time
    import torch
    import sparkdl.hvd as hvd
    from spark_deep_learning_b200.xgboost import gbdt
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(100 + r)
    X = torch.randn(rows, features, device=dev, generator=g)
    y = 2 * X[:, 0] + torch.sin(3 * X[:, 1]) + (X[:, 2] > 0.3).float() + 0.1 * torch.randn(rows, device=dev, generator=g)
    params = dict(n_estimators=rounds, max_depth=depth, learning_rate=0.3, use_gpu=True, objective="reg:squarederror")
    torch.cuda.synchronize()
    hvd.barrier()
    # the quantile sketch / binning is a one-time cost: measure it by training 1 round, then `rounds` more
    t0 = time.perf_counter()
    first, _ = gbdt.train(dict(params, n_estimators=1), X, y, device=dev)
    torch.cuda.synchronize()
    hvd.barrier()
    t1 = time.perf_counter()
    booster, _ = gbdt.train(params, X, y, device=dev)
    torch.cuda.synchronize()
    hvd.barrier()
    t2 = time.perf_counter()
    setup_plus_one = t1 - t0
    sec_per_round = max(1e-9, ((t2 - t1) - setup_plus_one)) / max(1, rounds - 1)
    # micro-benchmark of one tree level (64 nodes) with and without fusion
    dist = gbdt._Dist()
    cuts = gbdt.compute_cuts(X, float("nan"), 256, dist)
    bins = gbdt.bin_features(X, cuts, float("nan"))
    gh = torch.randint(-2 ** 20, 2 ** 20, (rows, 2), device=dev, dtype=torch.int32)
    pos = torch.randint(0, 32, (rows,), device=dev, dtype=torch.int32)
    res = {}
    for fused in (True, False):
        b = gbdt._HistBuilder(bins, dist, use_fused=fused)
        for _ in range(3):
            b.build(gh, pos, 32)
        torch.cuda.synchronize()
        hvd.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            b.build(gh, pos, 32)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100.0
        res["fused_us" if fused else "unfused_us"] = float(hvd.allreduce(torch.tensor([us], dtype=torch.float64), op=hvd.Max)[0])
    pred = booster.predict(X[:100000])
    rmse = float(((pred - y[:100000]) ** 2).mean().sqrt())
    t = hvd.allreduce(torch.tensor([sec_per_round], dtype=torch.float64), op=hvd.Max)
    hvd.comm_backend().check_error()
    out = {"np": n, "rows_per_worker": rows, "features": features, "depth": depth, "rounds": rounds,
           "sec_per_round": float(t[0]), "rows_per_sec": rows * n / float(t[0]), "rmse_after": rmse,
           "sketch_bin_plus_first_round_sec": setup_plus_one,
           "hist_level_fused_us": res["fused_us"], "hist_level_unfused_us": res["unfused_us"],
           "hist_bytes_per_level": 32 * features * 256 * 2 * 8}
    hvd.shutdown()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=1)
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--features", type=int, default=64)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import sparkdl
    res = sparkdl.HorovodRunner(np=-a.np).run(run, rows=a.rows, features=a.features, rounds=a.rounds)
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
