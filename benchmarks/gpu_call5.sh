#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -s 2>&1 | tail -40) > gpurun_out/pytest_gemm.log
cat gpurun_out/pytest_gemm.log
(timeout 600 python -m pytest tests/test_gpu_hvd.py -x -q -k "np1 or smoke" 2>&1 | tail -15) > gpurun_out/pytest_hvd1.log
tail -8 gpurun_out/pytest_hvd1.log
