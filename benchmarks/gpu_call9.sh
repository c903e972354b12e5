#!/bin/bash
# 8 GPUs: allreduce sweep vs NCCL, ResNet-50 and BERT-large scaling points (both arms)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
(timeout 600 python benchmarks/allreduce_sweep.py --np 8 --max-mb 1024 --iters 10 --out gpurun_out/sweep_np8.json 2>&1 | tail -70) > gpurun_out/sweep_np8.log
tail -3 gpurun_out/sweep_np8.log | cut -c1-600
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 400 $TR --master-port 29521 bench.py --gpus 8 --steps 15 --warmup 4 2> gpurun_out/bench8_ours.err | tail -1) > gpurun_out/bench8_ours.json
(timeout 400 $TR --master-port 29522 bench.py --impl reference --gpus 8 --steps 15 --warmup 4 2> gpurun_out/bench8_ref.err | tail -1) > gpurun_out/bench8_ref.json
(timeout 400 $TR --master-port 29523 bench.py --model bert-large --batch 64 --gpus 8 --steps 10 --warmup 3 2> gpurun_out/bert8_ours.err | tail -1) > gpurun_out/bert8_ours.json
(timeout 400 $TR --master-port 29524 bench.py --model bert-large --batch 64 --impl reference --gpus 8 --steps 10 --warmup 3 2> gpurun_out/bert8_ref.err | tail -1) > gpurun_out/bert8_ref.json
for f in bench8_ours bench8_ref bert8_ours bert8_ref; do python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('runtime'), d['clocks'])
except Exception as e: print('$f FAILED', e)
"; done
tail -5 gpurun_out/bench8_ours.err | cut -c1-300
