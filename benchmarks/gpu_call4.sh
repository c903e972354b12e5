#!/bin/bash
# 2 GPUs: full GPU test-suite (incl. multi-process VMM/multicast paths), torchrun external-gang bench
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_gpu4.log
tail -14 gpurun_out/pytest_gpu4.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 15 --warmup 4 2> gpurun_out/bench2_ref.err | tail -1) > gpurun_out/bench2_ref.json
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 15 --warmup 4 2> gpurun_out/bench2_ours.err | tail -1) > gpurun_out/bench2_ours.json
(timeout 600 python bench.py --gpus 1 --steps 15 --warmup 4 2> gpurun_out/bench1_ours.err | tail -1) > gpurun_out/bench1_ours.json
tail -4 gpurun_out/bench2_ours.err; cat gpurun_out/bench2_ref.json gpurun_out/bench2_ours.json gpurun_out/bench1_ours.json
