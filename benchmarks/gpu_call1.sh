#!/bin/bash
# first contact: capability matrix, GPU tests, smoke, 1-GPU bench (both arms), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpus.csv 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q -k "not multigpu" 2>&1 | tail -40) > gpurun_out/pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20) > gpurun_out/smoke.log
(timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_ref.err | tail -1) > gpurun_out/bench_ref.json
(timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-fused-model 2> gpurun_out/bench_ours.err | tail -1) > gpurun_out/bench_ours.json
tail -5 gpurun_out/bench_ours.err gpurun_out/bench_ref.err
cat gpurun_out/pytest_gpu.log | tail -15; cat gpurun_out/smoke.log; cat gpurun_out/bench_ref.json gpurun_out/bench_ours.json
