#!/bin/bash
# Parameterised GPU batch (replaces the one-shot gpu_callNN.sh scripts of round 1): every stage writes its own log
# under gpurun_out/ and runs under its own timeout.
#   gpurun --timeout 1500 -- 'bash benchmarks/gpu_batch.sh gemm pytest bench'
#   gpurun --gpus 2 --timeout 1500 -- 'bash benchmarks/gpu_batch.sh multi2 sweep2'
mkdir -p gpurun_out
for stage in "$@"; do
  echo "=== stage $stage ($(date +%T))"
  case "$stage" in
    gemm)      for g in misc qkv fwd bwd wgrad stats autograd; do
                 (timeout 150 python benchmarks/gemm_check.py $g 2>&1 | tail -80) > gpurun_out/gemm_check_$g.txt
                 echo "gemm_check $g: $(tail -1 gpurun_out/gemm_check_$g.txt) ($(grep -c '^BAD' gpurun_out/gemm_check_$g.txt) bad)"
               done ;;
    gemmbench) (timeout 400 python benchmarks/gemm_bench.py 2>&1 | tail -60) > gpurun_out/gemm_bench.txt; tail -50 gpurun_out/gemm_bench.txt ;;
    gemmquick) (timeout 200 python benchmarks/gemm_bench.py --quick 2>&1 | tail -30) > gpurun_out/gemm_bench_quick.txt; cat gpurun_out/gemm_bench_quick.txt ;;
    pytest)    (timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.txt; tail -6 gpurun_out/pytest_gpu.txt ;;
    pytestall) (timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest_gpu.txt; tail -30 gpurun_out/pytest_gpu.txt ;;
    multi)     (timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25) > gpurun_out/pytest_multi.txt; tail -12 gpurun_out/pytest_multi.txt ;;
    bench)     (timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench1.err | tail -1) > gpurun_out/bench1.json; cat gpurun_out/bench1.json ;;
    benchbert) (timeout 300 python bench.py --model bert-large --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bert1.err | tail -1) > gpurun_out/bert1.json; cat gpurun_out/bert1.json ;;
    benchref)  (timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/ref1.err | tail -1) > gpurun_out/ref1.json; cat gpurun_out/ref1.json ;;
    bertref)   (timeout 300 python bench.py --impl reference --model bert-large --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bertref1.err | tail -1) > gpurun_out/bertref1.json; cat gpurun_out/bertref1.json ;;
    fuzz1)     (timeout 300 python benchmarks/fuzz_collectives.py --np 1 --cases 200 2>&1 | tail -3) > gpurun_out/fuzz_np1.txt; cat gpurun_out/fuzz_np1.txt ;;
    fuzz2)     (timeout 400 python benchmarks/fuzz_collectives.py --np 2 --cases 300 2>&1 | tail -3) > gpurun_out/fuzz_np2.txt; cat gpurun_out/fuzz_np2.txt ;;
    fuzz8)     (timeout 400 python benchmarks/fuzz_collectives.py --np 8 --cases 150 2>&1 | tail -3) > gpurun_out/fuzz_np8.txt; cat gpurun_out/fuzz_np8.txt ;;
    probes)    (timeout 300 python benchmarks/resnet_probe.py 2>&1 | tail -30) > gpurun_out/resnet_probe.txt
               (timeout 300 python benchmarks/bert_probe.py 2>&1 | tail -30) > gpurun_out/bert_probe.txt
               tail -12 gpurun_out/resnet_probe.txt gpurun_out/bert_probe.txt ;;
    xgbprof)   (XGB_PROFILE=gpurun_out/xgb_profile.txt XGB_BENCH_SKLEARN=0 timeout 400 python benchmarks/xgboost_bench.py --np 1 2>&1 | tail -3) > gpurun_out/xgb_prof.log; cat gpurun_out/xgb_prof.log; cut -c1-200 gpurun_out/xgb_profile.txt | head -80 ;;
    xgb)       NPX=${XGB_NP:-1}; if [ -z "$XGB_SKIP_GA" ]; then (timeout 300 python -m pytest tests/test_gpu_xgboost.py -q -m gpu 2>&1 | tail -5) > gpurun_out/pytest_xgb.txt; cat gpurun_out/pytest_xgb.txt; fi
               (timeout 400 python benchmarks/xgboost_bench.py --np $NPX --out gpurun_out/xgboost_bench_np$NPX.json 2>&1 | tail -4) > gpurun_out/xgb_np$NPX.txt; cat gpurun_out/xgb_np$NPX.txt
               if [ -z "$XGB_SKIP_GA" ]; then (SPARKDL_GBDT_SMEM_HIST=0 timeout 400 python benchmarks/xgboost_bench.py --np $NPX --out gpurun_out/xgboost_bench_np${NPX}_globalatomics.json 2>&1 | tail -2) > gpurun_out/xgb_np${NPX}_ga.txt; cat gpurun_out/xgb_np${NPX}_ga.txt; fi ;;
    profbert)  (timeout 300 python benchmarks/step_profile.py bert 2>&1 | tail -5) > gpurun_out/step_profile_bert.log; cat gpurun_out/step_profile_bert.log ;;
    profresnet) (timeout 300 python benchmarks/step_profile.py resnet 2>&1 | tail -5) > gpurun_out/step_profile_resnet.log; cat gpurun_out/step_profile_resnet.log ;;
    sdpaprobe) (timeout 120 python benchmarks/sdpa_probe.py 2>&1 | tail -20) > gpurun_out/sdpa_probe.txt; cat gpurun_out/sdpa_probe.txt ;;
    bnvariants) for v in 0 1 2 3 4 5; do echo "variant $v"; (SPARKDL_BN_BWD_VARIANT=$v timeout 120 python benchmarks/bn_bench.py 2>&1 | grep bwd_apply); done > gpurun_out/bn_variants.txt; cat gpurun_out/bn_variants.txt ;;
    bnbench)   (timeout 300 python benchmarks/bn_bench.py 2>&1 | tail -40) > gpurun_out/bn_bench.txt; cat gpurun_out/bn_bench.txt ;;
    convbench) (timeout 400 python benchmarks/conv_bench.py 2>&1 | tail -30) > gpurun_out/conv_bench.txt; cat gpurun_out/conv_bench.txt ;;
    stembench) (timeout 300 python benchmarks/stem_bench.py 2>&1 | tail -14) > gpurun_out/stem_bench.txt; cat gpurun_out/stem_bench.txt ;;
    lntest)    (timeout 300 python -m pytest tests/test_gpu_layernorm.py -q -m gpu 2>&1 | tail -5) > gpurun_out/pytest_ln.txt; cat gpurun_out/pytest_ln.txt ;;
    capprobe)  (timeout 200 python benchmarks/capture_probe.py 2>&1 | grep -E "CAPTURE|eager|Error|error" | tail -20) > gpurun_out/capture_probe.txt; cat gpurun_out/capture_probe.txt
               (HID=1024 timeout 200 python benchmarks/capture_probe.py 2>&1 | grep -E "CAPTURE|eager|Error|error" | tail -20) > gpurun_out/capture_probe_1024.txt; cat gpurun_out/capture_probe_1024.txt ;;
    capprobe2) (timeout 200 python benchmarks/capture_probe2.py 2>&1 | grep -E "CAPTURE|Error|error" | tail -20) > gpurun_out/capture_probe2.txt; cat gpurun_out/capture_probe2.txt ;;
    sweep)     NPX=${SWEEP_NP:-2}; (timeout 900 python benchmarks/allreduce_sweep.py --np $NPX --max-mb ${SWEEP_MB:-256} --out gpurun_out/allreduce_sweep_np$NPX.json 2>&1 | tail -3) > gpurun_out/sweep_np$NPX.log; tail -2 gpurun_out/sweep_np$NPX.log ;;
    collsweep) NPX=${SWEEP_NP:-2}; (timeout 900 python benchmarks/collective_sweep.py --np $NPX --max-mb ${SWEEP_MB:-256} --out gpurun_out/collective_sweep_np$NPX.json 2>&1 | tail -3) > gpurun_out/collsweep_np$NPX.log; tail -2 gpurun_out/collsweep_np$NPX.log ;;
    bertprobe) (timeout 200 python benchmarks/bert_sink_probe.py 2>&1 | tail -40) > gpurun_out/bert_sink_probe.txt; tail -25 gpurun_out/bert_sink_probe.txt ;;
    gemmstaging) for d in 1 2; do (SPARKDL_GEMM_STAGING=$d timeout 120 python benchmarks/gemm_lab.py 2>&1 | sed "s/^dbg=0/staging=$d/" | tail -10); done > gpurun_out/gemm_staging.txt; cat gpurun_out/gemm_staging.txt ;;
    gemmlab)   for d in 0 1 2 3 4; do (SPARKDL_GEMM_DBG=$d timeout 120 python benchmarks/gemm_lab.py 2>&1 | grep -v cublas | tail -8); done > gpurun_out/gemm_lab.txt
               (timeout 120 python benchmarks/gemm_lab.py 2>&1 | grep cublas) >> gpurun_out/gemm_lab.txt; cat gpurun_out/gemm_lab.txt ;;
    gemmncu)   timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 6 -f -o gpurun_out/prof_gemm2 python benchmarks/gemm_lab.py ncu > gpurun_out/ncu_gemm2.log 2>&1; tail -3 gpurun_out/ncu_gemm2.log ;;
    ncuall)    # one --set full capture per hot kernel family (B200_PROFILING.md recipe); reports -> gpurun_out/r2_*.ncu-rep
               NCU="ncu --set full --clock-control none --import-source on -f"
               timeout 400 $NCU -k regex:gemm_kernel -s 3 -c 5 -o gpurun_out/r2_gemm python benchmarks/ncu_targets.py gemm > gpurun_out/ncu_r2_gemm.log 2>&1; tail -1 gpurun_out/ncu_r2_gemm.log
               timeout 300 $NCU -k regex:"bn_|maxpool|stem" -s 12 -c 6 -o gpurun_out/r2_bn python benchmarks/ncu_targets.py bn > gpurun_out/ncu_r2_bn.log 2>&1; tail -1 gpurun_out/ncu_r2_bn.log
               timeout 300 $NCU -k regex:"ln_" -s 6 -c 3 -o gpurun_out/r2_ln python benchmarks/ncu_targets.py ln > gpurun_out/ncu_r2_ln.log 2>&1; tail -1 gpurun_out/ncu_r2_ln.log
               timeout 300 $NCU -k regex:"maxpool" -s 4 -c 2 -o gpurun_out/r2_pool python benchmarks/ncu_targets.py pool > gpurun_out/ncu_r2_pool.log 2>&1; tail -1 gpurun_out/ncu_r2_pool.log
               timeout 300 $NCU -k regex:"xent" -s 2 -c 2 -o gpurun_out/r2_xent python benchmarks/ncu_targets.py xent > gpurun_out/ncu_r2_xent.log 2>&1; tail -1 gpurun_out/ncu_r2_xent.log
               timeout 300 $NCU -k regex:"stem_s2d" -s 1 -c 1 -o gpurun_out/r2_stem python benchmarks/ncu_targets.py stem > gpurun_out/ncu_r2_stem.log 2>&1; tail -1 gpurun_out/ncu_r2_stem.log
               timeout 300 $NCU -k regex:"gbdt_hist" -c 4 -o gpurun_out/r2_gbdt python benchmarks/ncu_targets.py gbdt > gpurun_out/ncu_r2_gbdt.log 2>&1; tail -1 gpurun_out/ncu_r2_gbdt.log
               timeout 300 $NCU -k regex:"fused_step" -s 3 -c 1 -o gpurun_out/r2_fused_step python benchmarks/ncu_targets.py fused_step > gpurun_out/ncu_r2_fused_step.log 2>&1; tail -1 gpurun_out/ncu_r2_fused_step.log
               ls -la gpurun_out/r2_*.ncu-rep ;;
    gemmlab2)  (timeout 300 python benchmarks/gemm_lab2.py 2>&1 | tail -60) > gpurun_out/gemm_lab2.txt; cat gpurun_out/gemm_lab2.txt ;;
    racedetail) (timeout 500 compute-sanitizer --print-limit 30 --tool racecheck python -m pytest tests/test_gpu_layernorm.py tests/test_gpu_xgboost.py -q -x -k "not distributed" 2>&1 | grep -E "=========|passed|failed" | grep -v "=========     at \|=========     in " | cut -c1-300 | tail -80) > gpurun_out/sanitizer_racecheck_ln_xent_gbdt.txt; cat gpurun_out/sanitizer_racecheck_ln_xent_gbdt.txt ;;
    nvlink)    NPX=${NVL_NP:-2}; (timeout 300 python benchmarks/nvlink_counters.py --np $NPX --out gpurun_out/nvlink_np$NPX.json 2>&1 | tail -12) > gpurun_out/nvlink_np$NPX.txt; cat gpurun_out/nvlink_np$NPX.txt ;;
    benchN)    N=${BENCH_NP:-2}; PORT=29611
               for m in resnet50 bert-large; do for impl in ${BENCH_IMPLS:-ours reference}; do
                 PORT=$((PORT+1)); tag=$(echo $m | cut -c1-4)${N}_$impl
                 (timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 --model $m --impl $impl 2> gpurun_out/$tag.err | tail -1) > gpurun_out/$tag.json; cut -c1-420 gpurun_out/$tag.json
               done; done ;;
    forklab)   for cfg in "0 0" "1 0" "1 1" "1 2" "1 3"; do set -- $cfg; echo "fork=$1 variant=$2"; (SPARKDL_BN_FORK=$1 SPARKDL_BN_FORK_VARIANT=$2 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 | grep -o '"ms_per_step": [0-9.]*'); done > gpurun_out/fork_lab.txt 2>&1; cat gpurun_out/fork_lab.txt ;;
    cudnnlab)  for lim in 10 0; do echo "cudnn benchmark_limit=$lim"; (SPARKDL_CUDNN_BENCH_LIMIT=$lim timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 | grep -o '"ms_per_step": [0-9.]*'); done > gpurun_out/cudnn_lab.txt 2>&1; cat gpurun_out/cudnn_lab.txt ;;
    sanitize)  bash benchmarks/sanitize.sh ;;
    *)         echo "unknown stage $stage" ;;
  esac
done
