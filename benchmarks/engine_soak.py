#!/usr/bin/env python
"""CPU soak runs of the negotiation engine through the public ``hvd`` API (multi-process over shared memory):

    python benchmarks/engine_soak.py shapes 2     # one name, alternating shapes / dtypes: cached responses are never reused
    python benchmarks/engine_soak.py mismatch 2   # a rank changes the shape of a cached tensor: error everywhere, engine usable after
    python benchmarks/engine_soak.py join 3       # uneven batch counts + join
    python benchmarks/engine_soak.py mixed 3      # 200 random allreduce / ragged allgather / broadcast / async groups

(The use-after-free of dropped async handles fixed in round 2 was found by a run of this kind.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparkdl  # noqa: E402

def main(mode):
    import faulthandler, torch
    faulthandler.enable()
    import sparkdl.hvd as hvd
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    res = []
    if mode == "shapes":
        # same name, shape / dtype alternate: a cached response must never be reused for a different signature
        for it in range(300):
            k = it % 4
            x = [torch.ones(8), torch.ones(16), torch.ones(8, dtype=torch.float64), torch.ones(2, 4)][k] * (r + 1)
            y = hvd.allreduce(x, op=hvd.Sum, name="same")
            exp = sum(range(1, n + 1))
            assert y.shape == x.shape and y.dtype == x.dtype and float(y.reshape(-1)[0]) == exp, (it, y)
        res = hvd.engine_stats()
    elif mode == "mismatch":
        # after the name is cached with shape [8], rank 1 submits shape [9]: must be an error on every rank, not a hang
        for it in range(5):
            hvd.allreduce(torch.ones(8), name="m")
        errs = 0
        try:
            hvd.allreduce(torch.ones(8 if r == 0 else 9), name="m")
        except hvd.HorovodInternalError as e:
            errs = 1
        ok = float(hvd.allreduce(torch.ones(8), name="m2")[0]) == 1.0     # engine still usable (average)
        res = (errs, ok)
    elif mode == "join":
        for it in range(50 + 10 * r):
            hvd.allreduce(torch.ones(4), name="j", op=hvd.Sum)
        last = hvd.join()
        res = last
    elif mode == "mixed":
        g = torch.Generator().manual_seed(7)
        for it in range(200):
            kind = int(torch.randint(0, 4, (1,), generator=g))
            sz = int(torch.randint(1, 2000, (1,), generator=g))
            if kind == 0:
                y = hvd.allreduce(torch.full((sz,), float(r)), op=hvd.Sum, name="x%d" % (it % 7))
                assert float(y[0]) == sum(range(n))
            elif kind == 1:
                y = hvd.allgather(torch.full((r + 1, 3), float(r)), name="g%d" % (it % 5))
                assert y.shape[0] == n * (n + 1) // 2
            elif kind == 2:
                y = hvd.broadcast(torch.full((sz,), float(r)), root_rank=it % n, name="b%d" % (it % 3))
                assert float(y[0]) == it % n
            else:
                hs = [hvd.allreduce_async(torch.ones(sz), name="as%d.%d" % (it % 2, j), op=hvd.Sum) for j in range(5)]
                for h in hs:
                    assert float(hvd.synchronize(h)[0]) == n
        res = hvd.engine_stats()
    return res

print(sparkdl.HorovodRunner(np=-int(sys.argv[2])).run(main, mode=sys.argv[1]))
