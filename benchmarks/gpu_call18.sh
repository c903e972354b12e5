#!/bin/bash
# final 8-GPU measurements
mkdir -p gpurun_out
(timeout 500 python benchmarks/allreduce_sweep.py --np 8 --max-mb 1024 --iters 10 --out gpurun_out/sweep_np8.json 2>&1 | tail -3) > gpurun_out/sweep_np8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 400 $TR --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/bench8_ours.err | tail -1) > gpurun_out/bench8_ours.json
(timeout 400 $TR --master-port 29542 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 2> gpurun_out/bench8_ref.err | tail -1) > gpurun_out/bench8_ref.json
(timeout 400 $TR --master-port 29543 bench.py --model bert-large --batch 64 --gpus 8 --steps 10 --warmup 3 2> gpurun_out/bert8_ours.err | tail -1) > gpurun_out/bert8_ours.json
(timeout 300 python benchmarks/xgboost_bench.py --np 8 --out gpurun_out/xgb_np8.json 2>&1 | tail -1 | cut -c1-500)
for f in bench8_ours bench8_ref bert8_ours; do python -c "
import json
try:
    d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('runtime'), d['clocks'])
except Exception as e: print('$f FAILED', e)
"; done
tail -2 gpurun_out/sweep_np8.log | cut -c1-400
grep -h "capture failed" gpurun_out/*8_ours.err | head -3 | cut -c1-300
