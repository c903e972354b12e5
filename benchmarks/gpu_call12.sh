#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu12.log
tail -8 gpurun_out/pytest_gpu12.log
(timeout 600 python benchmarks/xgboost_bench.py --np 2 --out gpurun_out/xgb_np2.json 2>&1 | tail -3) > gpurun_out/xgb_np2.log
tail -2 gpurun_out/xgb_np2.log | cut -c1-700
