#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_layernorm.py -q -x 2>&1 | tail -8) > gpurun_out/pytest_gpu17.log
tail -6 gpurun_out/pytest_gpu17.log
(timeout 400 python bench.py --model bert-large --batch 64 --gpus 1 --steps 10 --warmup 3 2> gpurun_out/g1_bert.err | tail -1) > gpurun_out/g1_bert.json
python -c "
import json
d=json.load(open('gpurun_out/g1_bert.json')); print('bert', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['runtime'], d['gpu_launches'], d['loss'])"
grep -h "capture failed" gpurun_out/g1_bert.err | head -2 | cut -c1-300
bash benchmarks/sanitize.sh
