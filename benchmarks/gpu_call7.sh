#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_xgboost.py -q 2>&1 | tail -12) > gpurun_out/pytest_gpu7.log
tail -6 gpurun_out/pytest_gpu7.log
(timeout 600 python benchmarks/bert_probe.py 2>&1 | tail -8) > gpurun_out/bert_probe.log
cat gpurun_out/bert_probe.log
head -48 gpurun_out/kernels_bert_ours.txt | cut -c1-200
head -30 gpurun_out/kernels_bert_hf.txt | cut -c1-200
