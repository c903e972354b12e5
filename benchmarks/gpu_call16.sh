#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_norm.py -q -x 2>&1 | tail -6) > gpurun_out/pytest_gpu16.log
tail -4 gpurun_out/pytest_gpu16.log
(timeout 400 python bench.py --gpus 1 --steps 15 --warmup 4 2> gpurun_out/g1_resnet.err | tail -1) > gpurun_out/g1_resnet.json
python -c "
import json
d=json.load(open('gpurun_out/g1_resnet.json')); print('resnet', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['runtime'], d['gpu_launches'], d['loss'])"
grep -h "capture failed" gpurun_out/g1_resnet.err | head -2 | cut -c1-300; tail -2 gpurun_out/g1_resnet.err | cut -c1-200
