#!/usr/bin/env python
"""NVLink byte counters around our fused compute+collective kernels at N GPUs (HorovodRunner(np=-N)).

ncu cannot replay a kernel that spins on flags written by a peer process, so the NVLink evidence is taken from the
driver's own counters instead: NVML's per-GPU cumulative NVLink data throughput (``NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX``
/ ``_RX``, KiB, all links) is read before and after K back-to-back launches of one kernel, device-timed with CUDA events
on the launching stream.  Reported per kernel: bytes that crossed this GPU's links per launch (max over ranks), the
algorithmic expectation, the achieved per-direction link rate and its fraction of the 770 GB/s measured peer-copy /
900 GB/s nominal figures of B200_PROFILING.md.

    python benchmarks/nvlink_counters.py --np 2 --out gpurun_out/nvlink_np2.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(iters=50):
    import ctypes
    import torch
    import pynvml
    import sparkdl.hvd as hvd
    from spark_deep_learning_b200.parallel import core

    hvd.init()
    rank, world = hvd.rank(), hvd.size()
    comm = hvd.comm_backend()
    lib = comm._lib
    dev = torch.device("cuda", torch.cuda.current_device())
    pynvml.nvmlInit()
    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    try:
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
    except Exception:  # pylint: disable=broad-except
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
    ALL = 0xFFFFFFFF

    def counters():
        out = {}
        for key, fid in (("tx", pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX), ("rx", pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX)):
            try:
                v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, ALL)])[0]
            except Exception:  # pylint: disable=broad-except
                v = pynvml.nvmlDeviceGetFieldValues(h, [fid])[0]
            if v.nvmlReturn != 0:
                return None
            out[key] = int(v.value.ullVal) * 1024      # KiB -> bytes
        return out

    stream = comm.channel_stream(2)
    tstream = torch.cuda.ExternalStream(stream, device=dev)
    rows = []

    def measure(name, fn, expect_bytes, note):
        with torch.cuda.stream(tstream):
            for _ in range(3):
                fn()
            tstream.synchronize()
            c0 = counters()          # NVML queries take milliseconds and differ per rank: re-align the ranks after them
            hvd.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(tstream)
            for _ in range(iters):
                fn()
            e1.record(tstream)
            tstream.synchronize()
            hvd.barrier()
            c1 = counters()
        us = e0.elapsed_time(e1) * 1e3 / iters
        if c0 is None or c1 is None:
            tx = rx = float("nan")
        else:
            tx, rx = (c1["tx"] - c0["tx"]) / iters, (c1["rx"] - c0["rx"]) / iters
        t = hvd.allreduce(torch.tensor([us, tx, rx], dtype=torch.float64), op=hvd.Max)
        us, tx, rx = float(t[0]), float(t[1]), float(t[2])
        gbs = max(tx, rx) / (us * 1e-6) / 1e9
        rows.append({"kernel": name, "np": world, "us": round(us, 2), "nvlink_tx_bytes_per_launch": round(tx),
                     "nvlink_rx_bytes_per_launch": round(rx), "expected_bytes_per_direction": expect_bytes,
                     "link_GBps_per_direction": round(gbs, 1), "frac_of_770": round(gbs / 770, 3),
                     "frac_of_900": round(gbs / 900, 3), "note": note})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)

    # ---- allreduce kernels on a symmetric buffer ------------------------------------------------------------------
    for mb in (32, 256):
        nb = mb << 20
        buf = comm.alloc_symmetric(nb)
        n = nb // 4
        buf.tensor(torch.float32, n).fill_(1.0)
        torch.cuda.synchronize()
        algos = [("allreduce_twoshot_kernel (P2P ld/st)", 2)]
        if comm.multicast and buf.multicast:
            algos.append(("allreduce_multimem_kernel (NVLS)", 3))
        for name, algo in algos:
            def fn(algo=algo, n=n, buf=buf):
                rc = lib.sdlc_allreduce(comm._h, 2, buf.id, 0, n, core.F32, 0, 1.0, 1.0 / world, algo, stream)
                assert rc >= 0, comm._err()
            if algo == 2:      # reduce-scatter pulls (N-1)/N x S from the peers, all-gather pushes/pulls the same again
                exp, note = 2 * (world - 1) * nb // world, "two-shot: 2 x (N-1)/N x S per direction"
            else:              # ld_reduce brings S/N per rank through the switch once, multimem.st sends S/N once
                exp, note = nb, "NVLS: this rank's S/N slice is reduced in the switch from N-1 peers and multicast back: ~S per direction"
            measure("%s fp32 %d MB" % (name, mb), fn, exp, note)
    # ---- fused reduce-scatter + AdamW + all-gather (what DistributedOptimizer(fused=True) launches per bucket) -------
    n = 32 << 20
    g, p = comm.alloc_symmetric(n * 2), comm.alloc_symmetric(n * 2)
    g.tensor(torch.bfloat16, n).fill_(0.01)
    p.tensor(torch.bfloat16, n).fill_(1.0)
    ns = lib.sdlc_fused_state_elems(comm._h, n, core.BF16)
    s1, s2, master = (torch.zeros(ns, device=dev) for _ in range(3))
    hyper = torch.tensor([2, 1e-3, 0.01, 0, 0, 0, 0.9, 0.999, 1e-8, 5.0] + [0.0] * 6, device=dev)
    torch.cuda.synchronize()

    def fstep():
        rc = lib.sdlc_fused_step(comm._h, 2, g.id, p.id, n, core.BF16, ctypes.c_void_p(s1.data_ptr()),
                                 ctypes.c_void_p(s2.data_ptr()), ctypes.c_void_p(master.data_ptr()), 1.0 / world,
                                 ctypes.c_void_p(hyper.data_ptr()), stream)
        assert rc == 0, comm._err()
    nb = n * 2
    measure("fused_step_kernel<bf16> AdamW 32M params", fstep, 2 * (world - 1) * nb // world,
            "gradient shard in ((N-1)/N x S) + updated parameters out ((N-1)/N x S)")
    comm.check_error()
    info = {"transport": comm.transport, "multicast": comm.multicast, "world": world,
            "counter": "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/RX (KiB, all links), delta over %d launches" % iters}
    hvd.shutdown()
    return {"info": info, "rows": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=2)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import sparkdl
    res = sparkdl.HorovodRunner(np=-a.np).run(run)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
