#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_norm.py tests/test_gpu_collectives.py -q -x -k "norm or wire or resnet" 2>&1 | tail -25) > gpurun_out/pytest_gpu3.log
tail -12 gpurun_out/pytest_gpu3.log
(timeout 900 python benchmarks/resnet_probe.py 2>&1 | tail -30) > gpurun_out/resnet_probe.log
tail -12 gpurun_out/resnet_probe.log
head -45 gpurun_out/kernels_torchvision.txt
head -45 gpurun_out/kernels_fused.txt
