#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
NCU="ncu --set full --clock-control none --import-source on"
(timeout 200 $NCU -k regex:gemm_bf16_kernel -s 4 -c 1 -f -o gpurun_out/prof_gemm python benchmarks/ncu_targets.py gemm 2>&1 | tail -1)
(timeout 200 $NCU -k regex:gemm_bf16_kernel -s 10 -c 1 -f -o gpurun_out/prof_gemm_bert python benchmarks/ncu_targets.py gemm 2>&1 | tail -1)
(timeout 200 $NCU -k regex:bn_ -s 12 -c 6 -f -o gpurun_out/prof_bn python benchmarks/ncu_targets.py bn 2>&1 | tail -1)
(timeout 200 $NCU -k regex:fused_step -s 2 -c 1 -f -o gpurun_out/prof_fused_step python benchmarks/ncu_targets.py fused_step 2>&1 | tail -1)
(timeout 200 $NCU -k "regex:ln_fwd|ln_bwd_kernel" -s 4 -c 2 -f -o gpurun_out/prof_ln python benchmarks/ncu_targets.py ln 2>&1 | tail -1)
(timeout 200 $NCU -k regex:maxpool -s 4 -c 2 -f -o gpurun_out/prof_pool python benchmarks/ncu_targets.py pool 2>&1 | tail -1)
ls gpurun_out/*.ncu-rep
