#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/pytest_gpu15.log
tail -5 gpurun_out/pytest_gpu15.log
