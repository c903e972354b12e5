#!/usr/bin/env python
"""Flagship benchmark (driver contract, see task statement).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--model resnet50|bert-large]

Metric (BASELINE.json): ResNet-50 images/sec, whole box, device-timed (CUDA events), max over
ranks, under ``HorovodRunner(np=-N)``; synthetic ``[B,3,224,224]`` data, random-init weights,
bf16 autocast.  N=1 goes through the real launcher (``HorovodRunner(np=-1)`` spawns the worker);
for N>1 the driver starts the ranks with torchrun, and ``HorovodRunner.run`` joins that gang
(``SPARKDL_EXTERNAL_GANG=1``).

``--impl reference`` runs the UNMODIFIED reference package from ``baseline/_ref`` through its own
public API (``sparkdl.HorovodRunner(np=-N).run(main)``, which executes ``main`` in-process); the
gradient exchange inside that main is torch DDP over NCCL - the NCCL-via-torch stand-in for
horovod+NCCL, since Horovod itself cannot be installed offline.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="resnet50", choices=["resnet50", "bert-large"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch size")
    ap.add_argument("--seq-len", type=int, default=128)
    ap.add_argument("--no-fused-optimizer", action="store_true")
    ap.add_argument("--no-fused-model", action="store_true")
    ap.add_argument("--wire", default=None, choices=[None, "bf16", "fp16", "fp8"])
    ap.add_argument("--no-cuda-graph", action="store_true")
    return ap.parse_args()


def main():
    args = parse()
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    n = args.gpus
    batch = args.batch or (256 if args.model == "resnet50" else 32)
    sys.path.insert(0, ROOT)
    from spark_deep_learning_b200.utils.clocks import ClockSampler  # harness only, not on the hot path

    if args.impl == "reference":
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        if not os.path.isdir(os.path.join(ref_dir, "sparkdl")):
            if rank == 0:
                print(json.dumps({"impl": "reference",
                                  "unavailable": "baseline/_ref/sparkdl missing (reference not installed)"}))
            return 0
        sys.path.insert(0, ref_dir)
        sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
        import sparkdl  # the reference package
        assert os.path.realpath(sparkdl.__file__).startswith(os.path.realpath(ref_dir)), sparkdl.__file__
        import reference_main
        fn, kwargs = reference_main.train, {}
    else:
        sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
        import cloudpickle
        import sparkdl
        import sparkdl_main
        cloudpickle.register_pickle_by_value(sparkdl_main)   # workers need not import benchmarks/
        fn = sparkdl_main.train
        kwargs = {"fused_optimizer": not args.no_fused_optimizer, "fused_model": not args.no_fused_model,
                  "wire": args.wire, "cuda_graph": not args.no_cuda_graph}
        if world > 1:
            os.environ["SPARKDL_EXTERNAL_GANG"] = "1"

    if n > 1 and world != n:
        raise SystemExit("bench.py --gpus %d must be launched with torchrun --nproc-per-node %d" % (n, n))
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))).start() if rank == 0 else None
    real_stdout = sys.stdout
    sys.stdout = sys.stderr            # worker log lines go to stderr; stdout carries ONE JSON line
    try:
        res = sparkdl.HorovodRunner(np=-n).run(
            fn, model_name="resnet50" if args.model == "resnet50" else "bert-large", batch=batch,
            steps=args.steps, warmup=args.warmup, seq_len=args.seq_len, **kwargs)
    finally:
        sys.stdout = real_stdout
    clocks = sampler.stop() if sampler is not None else None
    if rank != 0:
        return 0

    ms = res["dev_ms"] / res["steps"]
    global_batch = batch * n
    value = global_batch / (ms / 1e3)
    e2e = global_batch / (res["e2e_ms"] / res["e2e_steps"] / 1e3)
    unit = "images/sec" if args.model == "resnet50" else "samples/sec"
    out = {
        "metric": "ResNet-50 images/sec (whole box, device-timed, max over ranks) under HorovodRunner(np=-N)"
        if args.model == "resnet50" else
        "BERT-large samples/sec (whole box, device-timed, max over ranks) under HorovodRunner(np=-N)",
        "value": round(value, 2), "unit": unit, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init weights)",
        "impl": args.impl,
        "config": {"model": args.model, "global_batch": global_batch, "per_gpu_batch": batch,
                   "seq_len": args.seq_len if args.model != "resnet50" else None,
                   "parallelism": "dp%d" % n, "launcher": "HorovodRunner(np=-%d)" % n,
                   "l2": "inputs larger than L2 (fp32 [B,3,224,224] = %.0f MB per step; activations >> 126 MB)"
                   % (batch * 3 * 224 * 224 * 4 / 1e6) if args.model == "resnet50" else
                   "working set (weights+activations) >> 126 MB L2",
                   "optimizer": "SGD momentum 0.9 wd 1e-4" if args.model == "resnet50" else "AdamW"},
        "clocks": clocks,
        "e2e": {"value": round(e2e, 2), "unit": unit, "h2d_bytes_per_step": res["h2d_bytes"],
                "d2h_bytes_per_step": res["d2h_bytes"], "ms_per_step": round(res["e2e_ms"] / res["e2e_steps"], 3)},
        "gpu_launches": res.get("gpu_launches", 0),
        "loss": res.get("loss"),
    }
    if args.impl == "ours":
        out["runtime"] = {k: res.get(k) for k in ("transport", "multicast", "buckets", "cuda_graph", "h2d_gbps", "numa")}
    else:
        out["reference_note"] = ("unmodified sparkdl 2.2.0-db1 HorovodRunner (in-process run); gradient exchange = "
                                 "torch DDP over NCCL 2.28.9, the NCCL-via-torch stand-in for horovod+NCCL")
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
