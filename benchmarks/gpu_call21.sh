#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_hvd.py -q -x -k "resume" 2>&1 | tail -12) > gpurun_out/pytest_gpu21.log
tail -8 gpurun_out/pytest_gpu21.log
