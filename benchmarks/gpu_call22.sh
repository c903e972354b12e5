#!/bin/bash
mkdir -p gpurun_out
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2)
(timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 2> gpurun_out/final_resnet.err | tail -1) > gpurun_out/final_resnet.json
python -c "
import json
d=json.load(open('gpurun_out/final_resnet.json')); print('resnet', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['runtime'], d['gpu_launches'])"
