"""CUDA-graph capture of a whole training step (zero_grad + forward + backward + the fused
DistributedOptimizer kernels) - "CUDA streams and graphs instead of a tracing compiler".

A launch-bound step (BERT-large issues ~3000 kernels) replays as one graph launch.  The collective kernels of
``hvd.DistributedOptimizer(fused=True)`` are captured with the rest: their hyper-parameters live in device memory
(learning-rate changes: ``optimizer.sync_hyperparams()`` before the replay), the step counter advances on the device,
and the in-kernel peer barriers synchronise the replays of all ranks exactly like eager launches.
"""
import torch


class GraphedTrainStep(object):
    """``step = GraphedTrainStep(fn, (x, y)); loss = step(x, y)``.

    ``fn(*inputs) -> loss`` must run the complete step (``opt.zero_grad(); loss = ...; loss.backward(); opt.step()``)
    with static shapes.  The first ``warmup`` calls run eagerly on a side stream (allocations, cuDNN autotuning,
    lazy state of the optimizer), then the step is captured once and replayed."""

    def __init__(self, fn, example_inputs, warmup=3):
        self.fn = fn
        self.static_inputs = [torch.empty_like(t) if torch.is_tensor(t) else t for t in example_inputs]
        for s, t in zip(self.static_inputs, example_inputs):
            if torch.is_tensor(t):
                s.copy_(t)
        self.graph = None
        self.out = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.out = fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from spark_deep_learning_b200.utils import launch_counter
        comm = None
        try:
            from spark_deep_learning_b200.parallel import hvd
            comm = hvd.comm_backend() if hvd.is_initialized() else None
        except Exception:  # pylint: disable=broad-except
            comm = None
        self._comm = comm
        n0 = launch_counter.total(comm)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.out = fn(*self.static_inputs)
        self.graph = g
        # kernels of ours inside one replay (the Python-side counters do not run during replays)
        self.launches_per_replay = launch_counter.total(comm) - n0

    def __call__(self, *inputs):
        for s, t in zip(self.static_inputs, inputs):
            if torch.is_tensor(t) and t is not s:
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        if self._comm is not None:
            self._comm.check_error()   # in-kernel peer-wait timeouts surface here (host-mapped flag, no sync)
        return self.out
