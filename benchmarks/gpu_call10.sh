#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
(timeout 300 $NCU -k regex:gemm_bf16_kernel -s 4 -c 2 -f -o gpurun_out/prof_gemm python benchmarks/ncu_targets.py gemm 2>&1 | tail -3)
(timeout 300 $NCU -k regex:gemm_bf16_kernel -s 10 -c 1 -f -o gpurun_out/prof_gemm_bert python benchmarks/ncu_targets.py gemm 2>&1 | tail -2)
(timeout 300 $NCU -k regex:bn_ -s 12 -c 6 -f -o gpurun_out/prof_bn python benchmarks/ncu_targets.py bn 2>&1 | tail -3)
(timeout 300 $NCU -k regex:allreduce_twoshot -s 2 -c 1 -f -o gpurun_out/prof_allreduce python benchmarks/ncu_targets.py allreduce 2>&1 | tail -2)
(timeout 300 $NCU -k regex:fused_step -s 2 -c 1 -f -o gpurun_out/prof_fused_step python benchmarks/ncu_targets.py fused_step 2>&1 | tail -2)
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200)
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
