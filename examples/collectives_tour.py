#!/usr/bin/env python
"""A tour of the hvd.* calls inside a pickled main (works on CPU slots as well as GPUs).

    python examples/collectives_tour.py --np -2
"""
import argparse


def tour():
    import torch
    import sparkdl.hvd as hvd
    from sparkdl.horovod import log_to_driver
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    t = torch.full((4,), float(r + 1), device=dev)
    out = {
        "allreduce_avg": hvd.allreduce(t).tolist(),
        "allreduce_sum": hvd.allreduce(t, op=hvd.Sum).tolist(),
        "allgather_rows": hvd.allgather(torch.full((r + 1, 2), float(r), device=dev)).shape[0],
        "broadcast": hvd.broadcast(t, root_rank=n - 1).tolist(),
        "alltoall": hvd.alltoall(torch.arange(n, dtype=torch.float32, device=dev) + 10 * r).tolist(),
        "reducescatter": hvd.reducescatter(torch.ones(n, 3, device=dev), op=hvd.Sum).tolist(),
        "objects": hvd.allgather_object({"rank": r}),
    }
    handles = [hvd.allreduce_async(torch.ones(3, device=dev) * i, name="t%d" % i, op=hvd.Sum) for i in range(4)]
    out["async"] = [hvd.synchronize(h)[0].item() for h in handles]
    hvd.barrier()
    if r == 0:
        log_to_driver("tour finished on %d ranks" % n)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=-2)
    a = ap.parse_args()
    from sparkdl import HorovodRunner
    import json
    print(json.dumps(HorovodRunner(np=a.np).run(tour), indent=1))
