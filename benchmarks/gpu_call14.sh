#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_collectives.py tests/test_gpu_hvd.py tests/test_gpu_xgboost.py -q -x -k "wire or bf16 or fp8 or xgboost" 2>&1 | tail -6) > gpurun_out/pytest_gpu14.log
tail -4 gpurun_out/pytest_gpu14.log
(timeout 300 python benchmarks/xgboost_bench.py --np 2 --out gpurun_out/xgb_np2.json 2>&1 | tail -1 | cut -c1-600)
(timeout 600 python benchmarks/allreduce_sweep.py --np 2 --max-mb 256 --iters 10 --no-nccl --out gpurun_out/sweep_np2b.json 2>&1 | grep float32 | tail -6 | cut -c1-700)
