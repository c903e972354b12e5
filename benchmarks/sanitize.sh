#!/bin/bash
# compute-sanitizer pass over the kernels: memcheck on the GEMM / BN / pool / collective tests with
# virtual ranks.  Slow (10-50x); run with a tight selection.  Output: gpurun_out/sanitizer_*.txt
mkdir -p gpurun_out
CS="compute-sanitizer --error-exitcode 7 --print-limit 20"
(timeout 600 $CS --tool memcheck python -m pytest tests/test_gpu_collectives.py -q -x -k "allreduce_sum and 2-algo or pack or reducescatter or pull" 2>&1 | tail -15) > gpurun_out/sanitizer_memcheck_collectives.txt
(timeout 600 $CS --tool memcheck python -m pytest tests/test_gpu_norm.py -q -x -k "64-56-8 or maxpool" 2>&1 | tail -15) > gpurun_out/sanitizer_memcheck_norm.txt
tail -4 gpurun_out/sanitizer_memcheck_collectives.txt gpurun_out/sanitizer_memcheck_norm.txt
