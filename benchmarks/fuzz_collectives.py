#!/usr/bin/env python
"""Seeded fuzz of every collective (dtypes, zero-sized dims, ragged allgather, uneven alltoall, reducescatter
remainders) against locally recomputed expectations - the GPU twin of tests/test_hvd_cpu.py::test_collectives_fuzz_*.

    python benchmarks/fuzz_collectives.py --np 2 --cases 300 --seed 1 [--cpu]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
def main(cases, seed):
    import random
    import torch
    import sparkdl.hvd as hvd
    hvd.init()
    dev = torch.device('cuda', torch.cuda.current_device()) if hvd.comm_backend() is not None else torch.device('cpu')
    r, n = hvd.rank(), hvd.size()
    rnd = random.Random(seed)                      # same stream on every rank
    dtypes = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int32, torch.int64, torch.uint8]
    bad = []
    def data(shape, dt, rank):
        g = torch.Generator().manual_seed((rank * 1000003 + len(shape) * 7919 + sum(int(v) * (i + 1) for i, v in enumerate(shape)) * 31 + dtypes.index(dt)) % (2 ** 31))
        t = torch.randint(0, 5, tuple(shape), generator=g)
        return t.to(dt).to(dev)
    for c in range(cases):
        kind = rnd.choice(["allreduce", "allgather", "broadcast", "alltoall", "reducescatter"])
        dt = rnd.choice(dtypes)
        nd = rnd.randint(1, 3)
        shape = [rnd.choice([0, 1, 2, 3, 7, 33]) for _ in range(nd)]
        try:
            if kind == "allreduce":
                op = rnd.choice([hvd.Sum, hvd.Min, hvd.Max] + ([hvd.Average] if dt.is_floating_point else []))
                mine = data(shape, dt, r)
                out = hvd.allreduce(mine, op=op)
                allv = torch.stack([data(shape, dt, q).double() for q in range(n)])
                exp = {hvd.Sum: allv.sum(0), hvd.Min: allv.min(0).values if allv.numel() else allv.sum(0), hvd.Max: allv.max(0).values if allv.numel() else allv.sum(0), hvd.Average: allv.mean(0)}[op]
                ok = out.shape == mine.shape and out.dtype == dt and torch.allclose(out.double(), exp, rtol=2e-2, atol=1e-2)
            elif kind == "allgather":
                shapes = [[rnd.choice([0, 1, 4])] + shape[1:] for _ in range(n)]
                out = hvd.allgather(data(shapes[r], dt, r))
                exp = torch.cat([data(shapes[q], dt, q) for q in range(n)], 0)
                ok = out.shape == exp.shape and torch.equal(out, exp)
            elif kind == "broadcast":
                root = rnd.randrange(n)
                out = hvd.broadcast(data(shape, dt, r), root_rank=root)
                ok = torch.equal(out, data(shape, dt, root))
            elif kind == "alltoall":
                splits = [[rnd.choice([0, 1, 3]) for _ in range(n)] for _ in range(n)]      # [src][dst]
                sh = [sum(splits[r])] + shape[1:]
                mine = data(sh, dt, r)
                out, rs = hvd.alltoall(mine, splits=splits[r])
                parts = []
                for q in range(n):
                    src = data([sum(splits[q])] + shape[1:], dt, q)
                    off = sum(splits[q][:r])
                    parts.append(src[off:off + splits[q][r]])
                exp = torch.cat(parts, 0)
                ok = torch.equal(out, exp) and rs.cpu().tolist() == [splits[q][r] for q in range(n)]
            else:
                shape[0] = rnd.choice([0, 1, 2, 3, 5, 8])
                mine = data(shape, dt, r)
                out = hvd.reducescatter(mine, op=hvd.Sum)
                tot = torch.stack([data(shape, dt, q).double() for q in range(n)]).sum(0)
                d0 = shape[0]; rows = [d0 // n + (1 if q < d0 % n else 0) for q in range(n)]
                off = sum(rows[:r])
                ok = torch.allclose(out.double(), tot[off:off + rows[r]], rtol=2e-2, atol=1e-2) and out.shape[0] == rows[r]
        except Exception as e:
            ok = False
            bad.append((c, kind, str(dt), shape, "EXC " + repr(e)[:200]))
            break
        if not ok:
            bad.append((c, kind, str(dt), shape))
    return bad
if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=2)
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    if a.cpu:
        os.environ["SPARKDL_NUM_GPUS"] = "0"
        os.environ["SPARKDL_FORCE_CPU"] = "1"
    from sparkdl import HorovodRunner
    bad = HorovodRunner(np=-a.np).run(main, cases=a.cases, seed=a.seed)
    print("fuzz_collectives np=%d cases=%d seed=%d: %s" % (a.np, a.cases, a.seed, "OK" if not bad else "FAILED %r" % (bad[:5],)))
    sys.exit(0 if not bad else 1)
