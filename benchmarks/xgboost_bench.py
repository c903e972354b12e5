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


def run(rows=1000000, features=64, rounds=8, depth=6):
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
    stamps = []

    def stamp(_env):                            # per-round callback (public API): device-synchronised wall clock
        torch.cuda.synchronize()
        stamps.append(time.perf_counter())
        return False

    gbdt.train(dict(params, n_estimators=1), X, y, device=dev)      # warm-up: library load, symmetric buffers, allocator
    torch.cuda.synchronize()
    hvd.barrier()
    t0 = time.perf_counter()
    booster, _ = gbdt.train(params, X, y, device=dev, callbacks=[stamp])
    # sketch + binning + the first round are a one-time cost; the per-round cost is the median gap between callbacks
    setup_plus_one = stamps[0] - t0
    gaps = sorted(b_ - a_ for a_, b_ in zip(stamps[:-1], stamps[1:]))
    sec_per_round = gaps[len(gaps) // 2]
    if os.environ.get("XGB_PROFILE") and r == 0 and n == 1:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            gbdt.train(dict(params, n_estimators=2), X, y, device=dev)
            torch.cuda.synchronize()
        with open(os.environ["XGB_PROFILE"], "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
            f.write("\n\n")
            f.write(prof.key_averages().table(sort_by="cpu_time_total", row_limit=25, max_name_column_width=60))
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
    # comparison arm: the unfused histogram kernel followed by an NCCL allreduce of the same int64 bytes
    res["unfused_plus_nccl_us"] = None
    if n > 1:
        try:
            import torch.distributed as tdist
            if not tdist.is_initialized():
                tdist.init_process_group("nccl", device_id=dev)
            single = gbdt._Dist.__new__(gbdt._Dist)
            single.hvd, single.size, single.rank, single._n = None, 1, 0, 0
            b = gbdt._HistBuilder(bins, single, use_fused=False)
            for _ in range(3):
                tdist.all_reduce(b.build(gh, pos, 32))
            torch.cuda.synchronize()
            hvd.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                tdist.all_reduce(b.build(gh, pos, 32))
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100.0
            res["unfused_plus_nccl_us"] = float(hvd.allreduce(torch.tensor([us], dtype=torch.float64), op=hvd.Max)[0])
            tdist.destroy_process_group()
        except Exception as e:  # pylint: disable=broad-except
            res["unfused_plus_nccl_us"] = "unavailable: %r" % (e,)
    # context: scikit-learn's HistGradientBoosting (CPU, all cores) on rank 0's shard, same depth / bins / rounds
    hgb_sec_per_round = None
    if r == 0 and os.environ.get("XGB_BENCH_SKLEARN", "1") != "0":
        try:
            from sklearn.ensemble import HistGradientBoostingRegressor
            Xc, yc = X.cpu().numpy(), y.cpu().numpy()
            t0 = time.perf_counter()
            HistGradientBoostingRegressor(max_iter=rounds, max_depth=depth, max_leaf_nodes=None, learning_rate=0.3,
                                          max_bins=255, early_stopping=False).fit(Xc, yc)
            hgb_sec_per_round = (time.perf_counter() - t0) / rounds
        except Exception as e:  # pylint: disable=broad-except
            hgb_sec_per_round = "unavailable: %r" % (e,)
    pred = booster.predict(X[:100000])
    rmse = float(((pred - y[:100000]) ** 2).mean().sqrt())
    t = hvd.allreduce(torch.tensor([sec_per_round], dtype=torch.float64), op=hvd.Max)
    hvd.comm_backend().check_error()
    out = {"np": n, "rows_per_worker": rows, "features": features, "depth": depth, "rounds": rounds,
           "sec_per_round": float(t[0]), "rows_per_sec": rows * n / float(t[0]), "rmse_after": rmse,
           "sketch_bin_plus_first_round_sec": setup_plus_one,
           "hist_level_fused_us": res["fused_us"], "hist_level_unfused_us": res["unfused_us"],
           "hist_level_unfused_plus_nccl_us": res["unfused_plus_nccl_us"],
           "sklearn_hgb_sec_per_round_1shard_cpu": hgb_sec_per_round,
           "smem_hist": os.environ.get("SPARKDL_GBDT_SMEM_HIST", "1") != "0",
           "hist_bytes_per_level": 32 * features * 256 * 2 * 8}
    hvd.shutdown()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=1)
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--features", type=int, default=64)
    ap.add_argument("--rounds", type=int, default=8)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import sparkdl
    res = sparkdl.HorovodRunner(np=-a.np).run(run, rows=a.rows, features=a.features, rounds=a.rounds)
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
