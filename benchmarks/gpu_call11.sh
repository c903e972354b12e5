#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "not multigpu and not np1 and not smoke" -s 2>&1 | grep -E "TFLOP|passed|failed|Error|error|BAD" | tail -12) > gpurun_out/pytest_gpu11.log
cat gpurun_out/pytest_gpu11.log
(timeout 600 python benchmarks/bert_probe.py 2>&1 | grep "ms/step") > gpurun_out/bert_probe3.log
cat gpurun_out/bert_probe3.log
(timeout 600 python bench.py --model bert-large --batch 64 --gpus 1 --steps 10 --warmup 3 2> gpurun_out/bert1_ours.err | tail -1) > gpurun_out/bert1_ours.json
python -c "
import json; d=json.load(open('gpurun_out/bert1_ours.json')); print('bert ours', d['value'], d['ms_per_step'], d['e2e']['value'])"
(timeout 600 python bench.py --gpus 1 --steps 15 --warmup 4 2> gpurun_out/bench1_ours.err | tail -1) > gpurun_out/bench1_ours.json
python -c "
import json; d=json.load(open('gpurun_out/bench1_ours.json')); print('resnet ours', d['value'], d['ms_per_step'], d['e2e']['value'])"
tail -3 gpurun_out/bench1_ours.err | cut -c1-300
