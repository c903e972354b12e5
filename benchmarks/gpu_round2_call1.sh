#!/bin/bash
# First GPU call of the next round (1 GPU): re-validate everything written after the last round-1 call, refresh the
# per-kernel time tables of both flagship models, and answer the open BERT-glue questions in one go.
#   gpurun --timeout 1500 -- 'bash benchmarks/gpu_round2_call1.sh'
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2_pytest_gpu.txt
(timeout 200 python benchmarks/bert_glue_probe.py 2>&1 | tail -40) > gpurun_out/r2_bert_glue.txt
(timeout 300 python benchmarks/resnet_probe.py 2>&1 | tail -20) > gpurun_out/r2_resnet_probe.txt
(timeout 300 python benchmarks/bert_probe.py 2>&1 | tail -20) > gpurun_out/r2_bert_probe.txt
(timeout 200 python benchmarks/fuzz_collectives.py --np 1 --cases 200 2>&1 | tail -2) > gpurun_out/r2_fuzz_np1.txt
(timeout 300 python benchmarks/gemm_bench.py 2>&1 | tail -8) > gpurun_out/r2_gemm_bench.txt
(SPARKDL_GEMM_EPI16=1 timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -4) > gpurun_out/r2_epi16_tests.txt
(SPARKDL_GEMM_EPI16=1 timeout 300 python benchmarks/gemm_bench.py 2>&1 | tail -8) > gpurun_out/r2_gemm_bench_epi16.txt
(timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 2> gpurun_out/r2_bench1.err | tail -1) > gpurun_out/r2_bench1.json
tail -3 gpurun_out/r2_pytest_gpu.txt; cat gpurun_out/r2_bench1.json
