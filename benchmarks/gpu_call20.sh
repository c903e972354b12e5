#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_layernorm.py tests/test_gpu_hvd.py -q -x -k "masks_change or autotune or smoke" 2>&1 | tail -6) > gpurun_out/pytest_gpu20.log
tail -4 gpurun_out/pytest_gpu20.log
