"""The training function run by OUR arm of bench.py - the code a user of this framework writes.

Public API only: ``sparkdl.hvd`` (init / broadcast_parameters / DistributedOptimizer) and the
model zoo.  Gradient averaging + optimizer update run in the fused sm_100a peer-memory kernels.
"""


def train(model_name="resnet50", batch=256, steps=20, warmup=5, e2e_steps=None, seq_len=128,
          fused_optimizer=True, fused_model=True, wire=None, cuda_graph=True):
    import os
    import time

    import torch

    import sparkdl.hvd as hvd
    from spark_deep_learning_b200 import models

    if wire:
        os.environ["SPARKDL_WIRE_DTYPE"] = wire
    hvd.init()
    rank, world = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(1234)
    torch.backends.cudnn.benchmark = True
    if os.environ.get("SPARKDL_CUDNN_BENCH_LIMIT"):
        torch.backends.cudnn.benchmark_limit = int(os.environ["SPARKDL_CUDNN_BENCH_LIMIT"])   # 0 = try every algorithm
    comm = hvd.comm_backend()

    bf16_weights = bool(fused_optimizer and fused_model)     # fp32 masters live in the fused optimizer's state
    if model_name == "resnet50":
        model = models.resnet50(fused=fused_model, bf16_weights=bf16_weights).to(dev).to(memory_format=torch.channels_last)
        base = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        host_x = torch.randn(batch, 3, 224, 224).pin_memory()
        host_y = torch.randint(0, 1000, (batch,)).pin_memory()
    else:
        model = models.bert_large(fused=fused_model).to(dev)
        base = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
        host_x = torch.randint(0, 30522, (batch, seq_len)).pin_memory()
        host_y = torch.randint(0, 30522, (batch, seq_len)).pin_memory()
    opt = hvd.DistributedOptimizer(base, named_parameters=model.named_parameters(), fused=fused_optimizer)
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    crit = torch.nn.CrossEntropyLoss()

    def fwd(x, y):
        if model_name == "resnet50" and bf16_weights:
            return crit(model(x).float(), y)   # the stem kernel casts fp32 NCHW -> bf16 channels-last itself
        if model_name == "resnet50":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return crit(model(x.contiguous(memory_format=torch.channels_last)), y)
        return model(x, labels=y)          # bf16 parameters, fp32 master weights in the fused optimizer

    dx, dy = host_x.to(dev), host_y.to(dev)
    # diagnostic for the end-to-end number: pinned host -> device bandwidth of this rank's input buffer (untimed region)
    torch.cuda.synchronize()
    hvd.barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        dx.copy_(host_x, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbps = 3 * host_x.numel() * host_x.element_size() / (c0.elapsed_time(c1) * 1e-3) / 1e9
    from spark_deep_learning_b200.utils import affinity
    numa = affinity.bind_to_gpu(dev.index)

    def step(x, y):
        opt.zero_grad()
        loss = fwd(x, y)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        hvd.barrier()
        torch.cuda.synchronize()

    graphed = False
    if cuda_graph and fused_optimizer:
        # the whole step (incl. the peer-memory optimizer kernels) replays as one CUDA graph
        from spark_deep_learning_b200.utils.graph import GraphedTrainStep
        eager_step = step
        try:
            step = GraphedTrainStep(eager_step, (dx, dy), warmup=3)
            graphed = True
        except Exception as e:  # pylint: disable=broad-except
            print("CUDA graph capture failed, running eagerly: %r" % (e,), flush=True)
            torch.cuda.synchronize()
            step = eager_step
    for _ in range(warmup):
        step(dx, dy)
    barrier()
    from spark_deep_learning_b200.utils import launch_counter
    l0 = launch_counter.total(comm)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(dx, dy)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = launch_counter.total(comm) - l0
    if graphed:
        launches = step.launches_per_replay * steps

    # end to end through the public input pipeline: every step copies its inputs from pinned host
    # memory (prefetched one step ahead on a copy stream) and reads its loss back on the host
    from spark_deep_learning_b200.utils.data import AsyncScalar, DevicePrefetcher
    k = e2e_steps or steps
    for x, y in DevicePrefetcher([(host_x, host_y)] * 2, dev):
        step(x, y).item()
    barrier()
    reader = AsyncScalar()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for x, y in DevicePrefetcher([(host_x, host_y)] * k, dev):
        reader.push(step(x, y))
        reader.pop() if reader._n > 1 else None
    loss_val = reader.drain()[-1]
    e3.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = e2.elapsed_time(e3)
    if comm is not None:
        comm.check_error()

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64)
    t = hvd.allreduce(t, op=hvd.Max)
    res = {"dev_ms": t[0].item(), "e2e_ms": t[1].item(), "steps": steps, "e2e_steps": k,
           "h2d_bytes": host_x.numel() * host_x.element_size() + host_y.numel() * host_y.element_size(),
           "d2h_bytes": 4, "e2e_wall_ms": e2e_wall_ms, "loss": loss_val, "world": world, "batch": batch,
           "gpu_launches": int(launches),
           "transport": getattr(comm, "transport", None), "multicast": getattr(comm, "multicast", None),
           "buckets": len(opt.bucket_summary()), "cuda_graph": graphed,
           "h2d_gbps": round(float(hvd.allreduce(torch.tensor([h2d_gbps], dtype=torch.float64), op=hvd.Min)[0]), 1),
           "numa": numa}
    hvd.shutdown()
    return res
