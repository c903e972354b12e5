#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_collectives.py tests/test_gpu_hvd.py tests/test_gpu_xgboost.py -q -x -k "fused or optimizer or xgboost or smoke" 2>&1 | tail -8) > gpurun_out/pytest_gpu13.log
tail -5 gpurun_out/pytest_gpu13.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 400 python bench.py --gpus 1 --steps 15 --warmup 4 2> gpurun_out/g1_resnet.err | tail -1) > gpurun_out/g1_resnet.json
(timeout 400 python bench.py --model bert-large --batch 64 --gpus 1 --steps 10 --warmup 3 2> gpurun_out/g1_bert.err | tail -1) > gpurun_out/g1_bert.json
(timeout 400 $TR --master-port 29531 bench.py --gpus 2 --steps 15 --warmup 4 2> gpurun_out/g2_resnet.err | tail -1) > gpurun_out/g2_resnet.json
(timeout 400 $TR --master-port 29532 bench.py --model bert-large --batch 64 --gpus 2 --steps 10 --warmup 3 2> gpurun_out/g2_bert.err | tail -1) > gpurun_out/g2_bert.json
(timeout 300 python benchmarks/xgboost_bench.py --np 2 --out gpurun_out/xgb_np2.json 2>&1 | tail -1 | cut -c1-600)
for f in g1_resnet g1_bert g2_resnet g2_bert; do python -c "
import json
try:
    d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['runtime'], d['gpu_launches'])
except Exception as e: print('$f FAILED', e)
"; done
grep -h "graph capture failed" gpurun_out/g*.err | head -3 | cut -c1-400
tail -3 gpurun_out/g2_bert.err | cut -c1-300
