#!/usr/bin/env python
"""Allreduce bus-bandwidth sweep (BASELINE.json config 5): 1 KB - 1 GB at np=-2/-4/-8, our
peer-memory kernels (one-shot / two-shot P2P / two-shot NVLS / wire formats) next to NCCL 2.28.9
driven through torch.distributed (the NCCL-via-torch stand-in for horovod+NCCL).

    python benchmarks/allreduce_sweep.py --np 8 [--max-mb 1024] [--out profiles/allreduce_sweep_np8.json]

All times are CUDA-event device times on the launching stream, max over ranks.
busbw = 2 (N-1)/N * bytes / t  (NCCL convention), against 900 GB/s nominal / 770 GB/s measured peer copy.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sweep(max_mb=1024, iters=20, warmup=5, dtypes=("float32", "bfloat16"), with_nccl=True, wires=("bf16", "fp8")):
    import os
    import torch
    import sparkdl.hvd as hvd
    from spark_deep_learning_b200.parallel import core

    hvd.init()
    rank, world = hvd.rank(), hvd.size()
    comm = hvd.comm_backend()
    dev = torch.device("cuda", torch.cuda.current_device())
    lib = comm._lib
    nccl_ok = False
    if with_nccl and world > 1:
        try:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=dev)
            nccl_ok = True
        except Exception as e:  # pylint: disable=broad-except
            print("NCCL init failed:", e, flush=True)
    max_bytes = max_mb << 20
    buf = comm.alloc_symmetric(max_bytes)
    sizes = []
    s = 1024
    while s <= max_bytes:
        sizes.append(s)
        s *= 4 if s < (1 << 20) else 2
    stream = comm.channel_stream(2)
    tstream = torch.cuda.ExternalStream(stream, device=dev)
    rows = []

    def timed(fn, nb):
        k = iters if nb <= (64 << 20) else max(3, iters // 4)
        with torch.cuda.stream(tstream):
            for _ in range(warmup if nb <= (64 << 20) else 2):
                fn()
            tstream.synchronize()
            hvd.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(tstream)
            for _ in range(k):
                fn()
            e1.record(tstream)
            tstream.synchronize()
        ms = e0.elapsed_time(e1) / k
        t = hvd.allreduce(torch.tensor([ms], dtype=torch.float64), op=hvd.Max)
        return float(t[0])

    for dname in dtypes:
        dt = getattr(torch, dname)
        code = {"float32": core.F32, "bfloat16": core.BF16}[dname]
        esz = 4 if dname == "float32" else 2
        view = buf.tensor(dt, max_bytes // esz)
        view.fill_(1.0)
        torch.cuda.synchronize()
        if dname == "float32":
            for w in wires:
                comm._wire_buffer(buf, max_bytes // 4, w)   # collective allocation, outside timing
        for nb in sizes:
            n = nb // esz
            row = {"dtype": dname, "bytes": nb, "np": world}
            algos = [("oneshot", 1), ("twoshot_p2p", 2)]
            if nb <= (64 << 10) and world > 1:
                algos.insert(0, ("ll", 4))          # barrier-free flag-in-data kernel
            if comm.multicast and buf.multicast:
                algos.append(("nvls", 3))
            for name, algo in algos:
                if algo == 1 and nb > (8 << 20):
                    continue

                def fn(algo=algo, n=n):
                    rc = lib.sdlc_allreduce(comm._h, 2, buf.id, 0, n, code, 0, 1.0, 1.0 / world, algo, stream)
                    assert rc >= 0, comm._err()
                ms = timed(fn, nb)
                row[name + "_us"] = round(ms * 1e3, 2)
                row[name + "_busbw"] = round(2 * (world - 1) / world * nb / (ms * 1e-3) / 1e9, 2) if world > 1 else None
            if nb >= (4 << 20):    # how many CTAs does it take to saturate NVLink?  (default: SPARKDL_COMM_CTAS = 32)
                for ctas in ((16, 64, 128) if nb <= (64 << 20) else (64, 128)):
                    lib.sdlc_comm_set_max_blocks(comm._h, ctas)
                    for name, algo in algos[1:]:
                        def fnb(algo=algo, n=n):
                            rc = lib.sdlc_allreduce(comm._h, 2, buf.id, 0, n, code, 0, 1.0, 1.0 / world, algo, stream)
                            assert rc >= 0, comm._err()
                        ms = timed(fnb, nb)
                        row["%s_ctas%d_busbw" % (name, ctas)] = round(2 * (world - 1) / world * nb / (ms * 1e-3) / 1e9, 2) if world > 1 else None
                lib.sdlc_comm_set_max_blocks(comm._h, int(os.environ.get("SPARKDL_COMM_CTAS", "32")))
            if dname == "float32" and nb >= (1 << 16):
                for w in wires:
                    def fnw(w=w, n=n):
                        comm.allreduce_symmetric(buf, n, dt, op=core.OP_SUM, post=1.0 / world, wire=w, channel=2,
                                                 wait_current=False, stream=stream)
                    ms = timed(fnw, nb)
                    row["wire_%s_us" % w] = round(ms * 1e3, 2)
                    row["wire_%s_busbw_equiv" % w] = round(2 * (world - 1) / world * nb / (ms * 1e-3) / 1e9, 2) if world > 1 else None
            if nccl_ok:
                import torch.distributed as dist
                t = torch.ones(n, dtype=dt, device=dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k = iters if nb <= (64 << 20) else max(3, iters // 4)
                for _ in range(warmup):
                    dist.all_reduce(t)
                torch.cuda.synchronize()
                dist.barrier()
                e0.record()
                for _ in range(k):
                    dist.all_reduce(t)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / k
                ms = float(hvd.allreduce(torch.tensor([ms], dtype=torch.float64), op=hvd.Max)[0])
                row["nccl_us"] = round(ms * 1e3, 2)
                row["nccl_busbw"] = round(2 * (world - 1) / world * nb / (ms * 1e-3) / 1e9, 2)
                del t
            comm.check_error()
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
            view.fill_(1.0)   # keep values bounded across repeated in-place sums
            torch.cuda.synchronize()
    info = {"transport": comm.transport, "multicast": comm.multicast, "world": world,
            "topology": getattr(comm, "topology", None)}
    if nccl_ok:
        import torch.distributed as dist
        dist.destroy_process_group()
    hvd.shutdown()
    return {"info": info, "rows": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=2)
    ap.add_argument("--max-mb", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-nccl", action="store_true")
    a = ap.parse_args()
    import sparkdl
    res = sparkdl.HorovodRunner(np=-a.np).run(sweep, max_mb=a.max_mb, iters=a.iters, with_nccl=not a.no_nccl)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
    print(json.dumps(res["info"]))
