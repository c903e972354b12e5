#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_hvd.py tests/test_gpu_xgboost.py -q -k "np2 or xgboost or gpu_training or fused_hist" 2>&1 | tail -40) > gpurun_out/pytest_gpu6.log
tail -25 gpurun_out/pytest_gpu6.log
(timeout 600 python bench.py --model bert-large --batch 64 --gpus 1 --steps 10 --warmup 3 --impl reference 2> gpurun_out/bert1_ref.err | tail -1) > gpurun_out/bert1_ref.json
(timeout 600 python bench.py --model bert-large --batch 64 --gpus 1 --steps 10 --warmup 3 2> gpurun_out/bert1_ours.err | tail -1) > gpurun_out/bert1_ours.json
tail -6 gpurun_out/bert1_ours.err; tail -3 gpurun_out/bert1_ref.err; cat gpurun_out/bert1_ref.json gpurun_out/bert1_ours.json
