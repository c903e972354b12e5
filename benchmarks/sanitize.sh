#!/bin/bash
# compute-sanitizer pass over the kernels: memcheck on the GEMM / BN / pool / collective tests with
# virtual ranks.  Slow (10-50x); run with a tight selection.  Output: gpurun_out/sanitizer_*.txt
mkdir -p gpurun_out
CS="compute-sanitizer --error-exitcode 7 --print-limit 20"
(timeout 600 $CS --tool memcheck python -m pytest tests/test_gpu_collectives.py -q -x -k "allreduce_sum and 2-algo or pack or reducescatter or pull" 2>&1 | tail -15) > gpurun_out/sanitizer_memcheck_collectives.txt
(timeout 600 $CS --tool memcheck python -m pytest tests/test_gpu_norm.py -q -x -k "64-56-8 or maxpool" 2>&1 | tail -15) > gpurun_out/sanitizer_memcheck_norm.txt
for f in gpurun_out/sanitizer_memcheck_*.txt; do echo "== $f"; tail -n 4 $f; done
# round 2: racecheck (shared-memory hazards) + synccheck (barrier misuse) on the virtual-rank collective tests and on
# the shared-memory heavy kernels (LayerNorm, BN, cross-entropy, GBDT histogram, column sums); synccheck on the tcgen05 GEMM
for tool in racecheck synccheck; do
  (timeout 500 $CS --tool $tool python -m pytest tests/test_gpu_collectives.py -q -x -k "allreduce_sum and 2- and float32 or reducescatter or pack or adamw" 2>&1 | tail -12) > gpurun_out/sanitizer_${tool}_collectives.txt
  (timeout 500 $CS --tool $tool python -m pytest tests/test_gpu_layernorm.py tests/test_gpu_xgboost.py -q -x -k "not distributed" 2>&1 | grep -E "=========|passed|failed" | grep -v "^=========     at \|^=========     in " | cut -c1-260 | tail -60) > gpurun_out/sanitizer_${tool}_ln_xent_gbdt.txt
  (timeout 400 $CS --tool $tool python -m pytest tests/test_gpu_norm.py -q -x -k "64-56-8 or maxpool or stem" 2>&1 | tail -12) > gpurun_out/sanitizer_${tool}_norm.txt
done
(timeout 500 $CS --tool synccheck python benchmarks/gemm_check.py misc 2>&1 | tail -12) > gpurun_out/sanitizer_synccheck_gemm.txt
for f in gpurun_out/sanitizer_racecheck_*.txt gpurun_out/sanitizer_synccheck_*.txt; do echo "== $f"; tail -n 3 $f; done
