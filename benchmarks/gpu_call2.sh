#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/pytest_gpu2.log
(timeout 900 python benchmarks/allreduce_sweep.py --np 2 --max-mb 256 --out gpurun_out/sweep_np2.json 2>&1 | tail -60) > gpurun_out/sweep_np2.log
tail -25 gpurun_out/pytest_gpu2.log; tail -45 gpurun_out/sweep_np2.log
