#!/usr/bin/env python
"""allgather / broadcast / reducescatter sweep through the public ``hvd`` API (C++ engine -> peer-memory kernels) next to
NCCL via torch.distributed (the NCCL-via-torch stand-in for horovod+NCCL), BASELINE.json config 5.

    python benchmarks/collective_sweep.py --np 2 --out profiles/collective_sweep_np2.json

Timing: CUDA events around K operations issued back to back (async handles, one synchronize at the end), max over ranks.
busbw: allgather / reducescatter (N-1)/N * total bytes / t, broadcast bytes / t.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sweep(max_mb=256, iters=10):
    import torch
    import torch.distributed as dist
    import sparkdl.hvd as hvd
    hvd.init()
    rank, world = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    rows = []
    sizes = []
    s = 1024
    while s <= (max_mb << 20):
        sizes.append(s)
        s *= 4
    seq = [0]

    def timed(fn, k):
        for _ in range(2):
            fn(k)            # same number of operations in flight as the timed call: the allocator is warm
        torch.cuda.synchronize()
        hvd.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k
        return float(hvd.allreduce(torch.tensor([ms], dtype=torch.float64), op=hvd.Max, name="t.%d" % seq[0])[0])

    for nb in sizes:
        n = nb // 2
        k = iters if nb <= (16 << 20) else 3
        x = torch.ones(n, dtype=torch.bfloat16, device=dev)
        row = {"bytes": nb, "np": world}
        # ---- allgather: every rank contributes nb bytes
        out_ag = torch.empty(n * world, dtype=torch.bfloat16, device=dev)

        def ours_ag(kk=1):
            seq[0] += 1
            hs = [hvd.allgather_async(x, name="ag.%d.%d" % (seq[0], i)) for i in range(kk)]
            for h in hs:
                hvd.synchronize(h)

        def nccl_ag(kk=1):
            for _ in range(kk):
                dist.all_gather_into_tensor(out_ag, x)
        t = timed(ours_ag, k)
        row["allgather_ours_us"] = round(t * 1e3, 2)
        row["allgather_ours_busbw"] = round((world - 1) * nb / (t * 1e-3) / 1e9, 2)
        t = timed(nccl_ag, k)
        row["allgather_nccl_us"] = round(t * 1e3, 2)
        row["allgather_nccl_busbw"] = round((world - 1) * nb / (t * 1e-3) / 1e9, 2)
        # ---- broadcast of nb bytes from rank 0

        def ours_bc(kk=1):
            seq[0] += 1
            hs = [hvd.broadcast_async_(x, root_rank=0, name="bc.%d.%d" % (seq[0], i)) for i in range(kk)]
            for h in hs:
                hvd.synchronize(h)

        def nccl_bc(kk=1):
            for _ in range(kk):
                dist.broadcast(x, 0)
        t = timed(ours_bc, k)
        row["broadcast_ours_us"] = round(t * 1e3, 2)
        row["broadcast_ours_busbw"] = round(nb / (t * 1e-3) / 1e9, 2)
        t = timed(nccl_bc, k)
        row["broadcast_nccl_us"] = round(t * 1e3, 2)
        row["broadcast_nccl_busbw"] = round(nb / (t * 1e-3) / 1e9, 2)
        # ---- reducescatter of an nb-byte input (rows = world * 8 so that it splits evenly)
        if n % (world * 8) == 0:
            xr = x.view(world * 8, -1)
            out_rs = torch.empty(n // world, dtype=torch.bfloat16, device=dev)

            def ours_rs(kk=1):
                seq[0] += 1
                hs = [hvd.reducescatter_async(xr, op=hvd.Sum, name="rs.%d.%d" % (seq[0], i)) for i in range(kk)]
                for h in hs:
                    hvd.synchronize(h)

            def nccl_rs(kk=1):
                for _ in range(kk):
                    dist.reduce_scatter_tensor(out_rs, x)
            t = timed(ours_rs, k)
            row["reducescatter_ours_us"] = round(t * 1e3, 2)
            row["reducescatter_ours_busbw"] = round((world - 1) / world * nb / (t * 1e-3) / 1e9, 2)
            t = timed(nccl_rs, k)
            row["reducescatter_nccl_us"] = round(t * 1e3, 2)
            row["reducescatter_nccl_busbw"] = round((world - 1) / world * nb / (t * 1e-3) / 1e9, 2)
        hvd.comm_backend().check_error()
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del x, out_ag
    comm = hvd.comm_backend()
    info = {"transport": comm.transport, "multicast": comm.multicast, "world": world}
    dist.destroy_process_group()
    hvd.shutdown()
    return {"info": info, "rows": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=2)
    ap.add_argument("--max-mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import sparkdl
    res = sparkdl.HorovodRunner(np=-a.np).run(sweep, max_mb=a.max_mb, iters=a.iters)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
    print(json.dumps(res["info"]))
