"""T1/T3 - hvd frontend over the C++ engine + shm data plane, on CPU, multi-process
(BASELINE.json config 1)."""
import os

import pytest

from sparkdl import HorovodRunner
from spark_deep_learning_b200.parallel.launcher import WorkerFailure


def test_config1_np1_allreduce_1k():
    """BASELINE.json config 1: HorovodRunner(np=1).run(main): hvd.init + allreduce of a 1k tensor."""
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        t = torch.arange(1024, dtype=torch.float32)
        out = hvd.allreduce(t)
        return hvd.rank(), hvd.size(), bool(torch.equal(out, t))
    import os
    os.environ["SPARKDL_NUM_GPUS"] = "0"
    try:
        assert HorovodRunner(np=1).run(main) == (0, 1, True)
    finally:
        del os.environ["SPARKDL_NUM_GPUS"]


def _collectives_main():
    import torch
    import sparkdl.hvd as hvd
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    res = {}
    # allreduce: average / sum / min / max / product over dtypes
    for dt in (torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int32, torch.int64):
        t = (torch.arange(1000) % 7 + r + 1).to(dt)
        exp = sum((torch.arange(1000) % 7 + q + 1).to(torch.float64) for q in range(n))
        s = hvd.allreduce(t, op=hvd.Sum)
        res["sum_%s" % dt] = bool(torch.allclose(s.double(), exp, rtol=1e-2))
    t = torch.full((5,), float(r + 1))
    res["avg"] = bool(torch.allclose(hvd.allreduce(t), torch.full((5,), (n + 1) / 2.0)))
    res["min"] = hvd.allreduce(t, op=hvd.Min)[0].item() == 1.0
    res["max"] = hvd.allreduce(t, op=hvd.Max)[0].item() == float(n)
    res["prod"] = hvd.allreduce(torch.full((3,), 2.0), op=hvd.Product)[0].item() == 2.0 ** n
    res["prepost"] = bool(torch.allclose(
        hvd.allreduce(torch.ones(4), op=hvd.Sum, prescale_factor=0.5, postscale_factor=4.0),
        torch.full((4,), 2.0 * n)))
    # in-place + non-contiguous
    m = torch.ones(4, 6)[:, ::2]
    hvd.allreduce_(m, op=hvd.Sum)
    res["inplace_noncontig"] = bool(torch.equal(m, torch.full((4, 3), float(n))))
    # larger than the staging area (chunked path): SPARKDL_SHM_STAGE_BYTES=64KiB in the test
    big = torch.full((100003,), float(r))
    res["big"] = bool(torch.allclose(hvd.allreduce(big, op=hvd.Sum), torch.full((100003,), n * (n - 1) / 2.0)))
    # grouped (fused by the engine)
    outs = hvd.grouped_allreduce([torch.full((k + 1,), float(r)) for k in range(10)], op=hvd.Sum)
    res["grouped"] = all(o.shape[0] == k + 1 and o[0].item() == n * (n - 1) / 2.0 for k, o in enumerate(outs))
    # allgather with ragged first dim
    g = hvd.allgather(torch.full((r + 1, 2), float(r)))
    res["allgather"] = g.shape == (n * (n + 1) // 2, 2) and g[-1, 0].item() == float(n - 1)
    # broadcast
    b = torch.full((3,), float(r))
    hvd.broadcast_(b, root_rank=n - 1)
    res["broadcast"] = b[0].item() == float(n - 1)
    res["bcast_obj"] = hvd.broadcast_object({"r": r} if r == 0 else None, 0) == {"r": 0}
    res["gather_obj"] = hvd.allgather_object("rank%d" % r) == ["rank%d" % q for q in range(n)]
    # alltoall: rank r sends (d+1) rows of value r to rank d
    send = torch.cat([torch.full((d + 1, 2), float(r)) for d in range(n)])
    recv, rs = hvd.alltoall(send, splits=[d + 1 for d in range(n)])
    res["alltoall"] = recv.shape == (n * (r + 1), 2) and rs.tolist() == [r + 1] * n and \
        [recv[q * (r + 1), 0].item() for q in range(n)] == [float(q) for q in range(n)]
    # reducescatter (sum): 2n+1 rows -> remainder row goes to rank 0
    x = torch.arange((2 * n + 1) * 3, dtype=torch.float32).reshape(2 * n + 1, 3)
    rsx = hvd.reducescatter(x, op=hvd.Sum)
    rows = 2 + (1 if r == 0 else 0)
    start = 2 * r + (1 if r > 0 else 0)
    res["reducescatter"] = rsx.shape == (rows, 3) and bool(torch.equal(rsx, x[start:start + rows] * n))
    hvd.barrier()
    res["stats"] = hvd.engine_stats()["responses"] > 0
    return r, res


@pytest.mark.parametrize("n", [2, 4])
def test_collectives(n, monkeypatch):
    monkeypatch.setenv("SPARKDL_SHM_STAGE_BYTES", str(64 << 10))
    r, res = HorovodRunner(np=-n).run(_collectives_main)
    bad = [k for k, v in res.items() if not v]
    assert r == 0 and not bad, bad


def test_out_of_order_submission_and_fusion():
    """Ranks submit named tensors in different orders; negotiation must still match them."""
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        names = ["t%d" % i for i in range(12)]
        order = names if r % 2 == 0 else list(reversed(names))
        hs = {nm: hvd.allreduce_async(torch.full((3,), float(int(nm[1:]) + r)), name=nm, op=hvd.Sum)
              for nm in order}
        ok = all(hvd.synchronize(hs[nm])[0].item() == n * int(nm[1:]) + n * (n - 1) / 2.0 for nm in names)
        return ok
    assert HorovodRunner(np=-3).run(main)


def test_mismatch_errors_on_all_ranks():
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        errs = []
        try:
            hvd.allreduce(torch.ones(3 + hvd.rank()), name="bad_shape")
        except hvd.HorovodInternalError as e:
            errs.append("shape" if "mismatched shapes" in str(e) else str(e))
        try:
            hvd.allreduce(torch.ones(3, dtype=torch.float32 if hvd.rank() else torch.float64), name="bad_dtype")
        except hvd.HorovodInternalError as e:
            errs.append("dtype" if "data types" in str(e) else str(e))
        # engine still usable afterwards
        ok = hvd.allreduce(torch.ones(2), op=hvd.Sum)[0].item() == hvd.size()
        both = hvd.allgather_object(errs)
        return ok, both
    ok, both = HorovodRunner(np=-2).run(main)
    assert ok and both == [["shape", "dtype"], ["shape", "dtype"]]


def test_duplicate_name_is_an_error():
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        h1 = hvd.allreduce_async(torch.ones(200000), name="dup")
        h2 = hvd.allreduce_async(torch.ones(200000), name="dup")
        msg = ""
        for h in (h1, h2):
            try:
                hvd.synchronize(h)
            except hvd.HorovodInternalError as e:
                msg = str(e)
        return msg
    msg = HorovodRunner(np=-2).run(main)
    # Timing dependent, as in Horovod (the check is local to each rank): the second submission may start after the
    # first finished (no error), be rejected here, or be rejected on the peer only - then this rank's second
    # operation can never be matched and is aborted when the peer leaves.
    assert "duplicate tensor name" in msg or "job aborted" in msg or msg == ""


def test_join_uneven_batches():
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        sums = []
        for step in range(r + 1):           # rank r has r+1 batches
            sums.append(hvd.allreduce(torch.ones(2), op=hvd.Sum, name="s%d" % step)[0].item())
        last = hvd.join()
        return r, sums, last
    r, sums, last = HorovodRunner(np=-3).run(main)
    assert sums == [3.0] and last == 2      # rank 0 only sees step 0, contributed to by all three


def test_distributed_optimizer_matches_single_process():
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 2))
        ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 2))
        ref.load_state_dict(model.state_dict())
        opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9),
                                       named_parameters=model.named_parameters(), bucket_bytes=256)
        ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
        hvd.broadcast_parameters(model.state_dict(), root_rank=0)
        hvd.broadcast_optimizer_state(opt, root_rank=0)
        g = torch.Generator().manual_seed(123)
        X = torch.randn(4 * n, 8, generator=g)
        Y = torch.randn(4 * n, 2, generator=g)
        for _ in range(3):
            opt.zero_grad()
            loss = ((model(X[4 * r:4 * r + 4]) - Y[4 * r:4 * r + 4]) ** 2).mean()
            loss.backward()
            opt.step()
            ropt.zero_grad()
            ((ref(X) - Y) ** 2).mean().backward()
            ropt.step()
        diff = max((a - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
        return diff, len(opt.bucket_summary())
    diff, nb = HorovodRunner(np=-2).run(main)
    assert diff < 1e-5 and nb >= 2


def test_distributed_optimizer_keeps_channels_last_weights_inside_the_buckets():
    """Bucket views keep a parameter's memory layout: a channels-last convolution weight stays channels-last (dense
    gradient accumulation, no per-step transposes) and training matches a single-process replica."""
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        torch.manual_seed(0)

        def make():
            return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 1),
                                       torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(4, 2))
        model, ref = make().to(memory_format=torch.channels_last), make()
        ref.load_state_dict(model.state_dict())
        w = model[0].weight
        assert not w.is_contiguous() and w.is_contiguous(memory_format=torch.channels_last)
        opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9),
                                       named_parameters=model.named_parameters(), bucket_bytes=512)
        ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
        hvd.broadcast_parameters(model.state_dict(), root_rank=0)
        layout = (w.grad.stride() == w.stride() and w.grad.is_contiguous(memory_format=torch.channels_last)
                  and model[2].weight.grad.is_contiguous())           # 1x1 kernel: both layouts coincide
        g = torch.Generator().manual_seed(5)
        X = torch.randn(4 * n, 3, 6, 6, generator=g)
        Y = torch.randn(4 * n, 2, generator=g)
        for _ in range(3):
            opt.zero_grad()
            ((model(X[4 * r:4 * r + 4].contiguous(memory_format=torch.channels_last)) - Y[4 * r:4 * r + 4]) ** 2).mean().backward()
            opt.step()
            ropt.zero_grad()
            ((ref(X) - Y) ** 2).mean().backward()
            ropt.step()
        diff = max((a - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
        return diff, layout, w.grad.stride() == w.stride()
    diff, layout, still = HorovodRunner(np=-2).run(main)
    assert layout and still and diff < 1e-5, (diff, layout, still)


def test_cpu_affinity_helper_parses_sysfs_lists_and_never_raises():
    from spark_deep_learning_b200.utils import affinity
    assert affinity._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist("") == set()
    info = affinity.bind_to_gpu(0)                   # no GPU here: nothing is bound, nothing raises
    assert info["bound"] is False and affinity.bind_to_gpu(0) is info


def test_dropped_async_handles_keep_their_tensors_until_the_engine_is_done():
    """An async handle that is dropped without synchronize() must not free tensors the engine thread still writes
    (a soak run used to corrupt the heap).  The frontend keeps them in an in-flight registry until completion."""
    def main():
        import gc
        import time
        import torch
        import sparkdl.hvd as hvd
        from spark_deep_learning_b200.parallel import hvd as H
        hvd.init()
        tracked = 0
        for rep in range(40):                                   # 40 x 32 operations whose handles are dropped at once
            for i in range(32):
                hvd.allreduce_async(torch.full((4096,), float(i)), name="drop.%d.%d" % (rep, i))
            tracked = max(tracked, len(H._S.inflight))          # the registry, not the (dead) handles, owns the tensors
            gc.collect()
        ok = hvd.allreduce(torch.ones(3), op=hvd.Sum, name="after")[0].item() == hvd.size()
        hvd.barrier()
        deadline = time.time() + 20
        while H._S.inflight and time.time() < deadline:
            H._purge_inflight()
            time.sleep(0.01)
        return tracked, len(H._S.inflight), ok
    tracked, left, ok = HorovodRunner(np=-2).run(main)
    assert ok and tracked >= 32 and left == 0, (tracked, left, ok)


def test_single_process_script_without_shutdown_leaves_no_shared_memory_behind():
    import glob
    import subprocess
    import sys
    before = set(glob.glob("/dev/shm/sparkdl-single*"))
    code = ("import torch, sparkdl.hvd as hvd; hvd.init(); "
            "print(float(hvd.allreduce(torch.ones(3), op=hvd.Sum)[0]))")          # no hvd.shutdown()
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and r.stdout.strip() == "1.0", r.stderr[-500:]
    assert set(glob.glob("/dev/shm/sparkdl-single*")) - before == set()


def test_dead_peer_fails_job_quickly():
    def main():
        import os
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        if hvd.rank() == 1:
            os._exit(9)
        hvd.allreduce(torch.ones(4))      # would hang forever without gang teardown
    with pytest.raises(WorkerFailure):
        HorovodRunner(np=-2).run(main)


def test_long_tensor_names_stay_unique():
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        base = "model." + "very_long_submodule_name." * 8           # > 127 bytes, common 127-byte prefix
        h1 = hvd.allreduce_async(torch.full((2,), 1.0), name=base + "weight", op=hvd.Sum)
        h2 = hvd.allreduce_async(torch.full((2,), 10.0), name=base + "bias", op=hvd.Sum)
        return hvd.synchronize(h1)[0].item(), hvd.synchronize(h2)[0].item()
    assert HorovodRunner(np=-2).run(main) == (2.0, 20.0)


def test_grouped_async_variants_and_elastic_state():
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        res = {}
        h = hvd.grouped_allgather_async([torch.full((r + 1, k + 1), float(r)) for k in range(3)])
        import time
        t0 = time.time()
        while not hvd.poll(h) and time.time() - t0 < 30:        # a grouped handle polls its members
            time.sleep(0.001)
        res["poll_group"] = hvd.poll(h)
        outs = hvd.synchronize(h)
        res["poll_done"] = hvd.poll(h)
        res["gag"] = [tuple(o.shape) for o in outs] == [(n * (n + 1) // 2, k + 1) for k in range(3)]
        h = hvd.grouped_reducescatter_async([torch.ones(n * 2, 3) * (k + 1) for k in range(2)], op=hvd.Sum)
        outs = hvd.synchronize(h)
        res["grs"] = all(tuple(o.shape) == (2, 3) and o[0, 0].item() == n * (k + 1) for k, o in enumerate(outs))

        # elastic state: sync from rank 0, commit, diverge, restore
        torch.manual_seed(100 + r)                                  # ranks start different on purpose
        model = torch.nn.Linear(4, 2)
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
        state = hvd.elastic.TorchState(model, opt, epoch=r * 10, batch=0)
        resets = []
        state.register_reset_callbacks([lambda: resets.append(1)])

        @hvd.elastic.run
        def train(st, fail):
            w0 = hvd.allgather(st.model.weight.detach().reshape(1, -1))
            same = bool(torch.allclose(w0[0], w0[-1])) and st.epoch == 0       # synced to rank 0's values
            st.epoch, st.batch = 3, 7
            st.commit()
            with torch.no_grad():
                st.model.weight.add_(1.0)
            st.epoch = 99
            if fail:
                raise hvd.HorovodInternalError("injected")
            return same

        res["synced"] = train(state, False)
        w_committed = state._model_saved["weight"].clone()           # 2nd call commits this + 1 (1st call's drift)
        try:
            train(state, True)
            res["raised"] = False
        except hvd.HorovodInternalError:
            res["raised"] = True
        res["rolled_back"] = state.epoch == 3 and state.batch == 7 and bool(torch.allclose(model.weight.detach(), w_committed + 1.0)) \
            and len(resets) == 1
        return res
    res = HorovodRunner(np=-2).run(main)
    assert all(res.values()), res


def test_fused_optimizer_request_on_a_cpu_gang_still_trains():
    """fused=True has no CPU implementation: it must degrade to allreduce + torch step, never to a silent no-op."""
    def main():
        import warnings
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        torch.manual_seed(0)
        model = torch.nn.Linear(6, 3)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.5), named_parameters=model.named_parameters(),
                                           fused=True)
        before = model.weight.detach().clone()
        x = torch.randn(8, 6, generator=torch.Generator().manual_seed(hvd.rank()))
        opt.zero_grad()
        model(x).pow(2).mean().backward()
        opt.step()
        moved = (model.weight.detach() - before).abs().max().item()
        same = hvd.allgather(model.weight.detach().reshape(1, -1))
        return moved, bool(torch.allclose(same[0], same[-1])), any("falling back" in str(x.message) for x in w)
    moved, same, warned = HorovodRunner(np=-2).run(main)
    assert moved > 1e-4 and same and warned


def test_sync_batch_norm_matches_full_batch_batch_norm():
    """hvd.SyncBatchNorm over 2 workers x half a batch == nn.BatchNorm2d over the whole batch: outputs, input gradients,
    parameter gradients (after the sum over workers) and running statistics."""
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        g = torch.Generator().manual_seed(0)
        X = torch.randn(4 * n, 3, 5, 5, generator=g, dtype=torch.float64)
        W = torch.randn(4 * n, 3, 5, 5, generator=g, dtype=torch.float64)      # loss = sum(y * W)
        ref = torch.nn.BatchNorm2d(3).double()
        with torch.no_grad():
            ref.weight.copy_(torch.tensor([1.5, 0.5, -1.0]))
            ref.bias.copy_(torch.tensor([0.1, -0.2, 0.3]))
        sbn = hvd.SyncBatchNorm(3).double()
        sbn.load_state_dict(ref.state_dict())
        xr = X.clone().requires_grad_()
        yr = ref(xr)                                    # exactly one training-mode forward (running stats update once)
        (yr * W).sum().backward()
        xl = X[4 * r:4 * r + 4].clone().requires_grad_()
        yl = sbn(xl)
        (yl * W[4 * r:4 * r + 4]).sum().backward()
        wg = hvd.allreduce(sbn.weight.grad, op=hvd.Sum)
        bg = hvd.allreduce(sbn.bias.grad, op=hvd.Sum)
        return {
            "y": bool(torch.allclose(yl, yr[4 * r:4 * r + 4].detach(), atol=1e-10)),
            "dx": bool(torch.allclose(xl.grad, xr.grad[4 * r:4 * r + 4], atol=1e-10)),
            "dw": bool(torch.allclose(wg, ref.weight.grad, atol=1e-9)) and bool(torch.allclose(bg, ref.bias.grad, atol=1e-9)),
            "rm": bool(torch.allclose(sbn.running_mean, ref.running_mean, atol=1e-10)),
            "rv": bool(torch.allclose(sbn.running_var, ref.running_var, atol=1e-10)),
        }
    res = HorovodRunner(np=-2).run(main)
    assert all(res.values()), res


def test_optimizer_recipes_and_differentiable_collectives():
    """Gradient accumulation (backward_passes_per_step), the clipping recipe (synchronize + skip_synchronize),
    gradient_predivide_factor, compression, in-place grouped allreduce, process-set stubs, Adasum refusal and the
    autograd rules of allreduce / allgather / broadcast."""
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        res = {}

        def make():
            torch.manual_seed(1)
            return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 2))
        gen = torch.Generator().manual_seed(9)
        X, Y = torch.randn(8 * n, 5, generator=gen), torch.randn(8 * n, 2, generator=gen)

        # (1) accumulation over 2 micro-batches + predivide: equals one big-batch SGD step on the global mean loss
        model, ref = make(), make()
        opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.2), named_parameters=model.named_parameters(),
                                       backward_passes_per_step=2, gradient_predivide_factor=2.0, bucket_bytes=64)
        opt.zero_grad()
        for k in range(2):
            lo = 8 * r + 4 * k
            ((model(X[lo:lo + 4]) - Y[lo:lo + 4]) ** 2).mean().div(2).backward()
        opt.step()
        ropt = torch.optim.SGD(ref.parameters(), lr=0.2)
        ((ref(X) - Y) ** 2).mean().backward()
        ropt.step()
        res["accumulate"] = max((a - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters())) < 1e-6

        # (2) clipping recipe: synchronize, clip the *averaged* gradients, step without a second reduction
        model, ref = make(), make()
        opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.2), named_parameters=model.named_parameters())
        opt.zero_grad()
        ((model(X[8 * r:8 * r + 8]) - Y[8 * r:8 * r + 8]) ** 2).mean().backward()
        opt.synchronize()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.05)
        with opt.skip_synchronize():
            opt.step()
        ropt = torch.optim.SGD(ref.parameters(), lr=0.2)
        ((ref(X) - Y) ** 2).mean().backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        ropt.step()
        res["clip"] = max((a - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters())) < 1e-6

        # (3) compression, in-place grouped allreduce
        t = torch.full((6,), 1.0 + r)
        c = hvd.allreduce(t, op=hvd.Sum, compression=hvd.Compression.fp16)
        res["compression"] = c.dtype == torch.float32 and bool(torch.allclose(c, torch.full((6,), n * (n + 1) / 2.0)))
        ts = [torch.full((k + 2,), float(r)) for k in range(3)]
        outs = hvd.grouped_allreduce_(ts, op=hvd.Sum)
        res["grouped_inplace"] = all(o is t_ and t_[0].item() == n * (n - 1) / 2.0 for o, t_ in zip(outs, ts))

        # (4) autograd through collectives: d/dx of sum over ranks
        x = torch.full((3,), float(r + 1), requires_grad=True)
        hvd.allreduce(x * x, op=hvd.Sum).sum().backward()           # every rank differentiates the same global sum
        res["d_allreduce"] = bool(torch.allclose(x.grad, 2 * x.detach() * n))
        x = torch.full((r + 1, 2), 1.0, requires_grad=True)
        g = hvd.allgather(x)
        (g * torch.arange(g.shape[0], dtype=torch.float32).view(-1, 1)).sum().backward()
        off = r * (r + 1) // 2
        want = torch.arange(off, off + r + 1, dtype=torch.float32).view(-1, 1).expand(-1, 2) * n
        res["d_allgather"] = bool(torch.allclose(x.grad, want))
        x = torch.full((2,), float(r + 1), requires_grad=True)
        (hvd.broadcast(x, root_rank=0) * (r + 1)).sum().backward()
        res["d_broadcast"] = bool(torch.allclose(x.grad, torch.full((2,), n * (n + 1) / 2.0) if r == 0 else torch.zeros(2)))

        # (5) refusals / stubs
        try:
            hvd.allreduce(torch.ones(2), op=hvd.Adasum)
            res["adasum"] = False
        except NotImplementedError:
            res["adasum"] = True
        res["global_set"] = hvd.add_process_set(list(range(n))) is hvd.global_process_set and hvd.global_process_set.size() == n
        try:
            hvd.add_process_set([0])
            res["subset"] = n == 1
        except NotImplementedError:
            res["subset"] = True
        return res
    res = HorovodRunner(np=-2).run(main)
    assert all(res.values()), res


def test_grad_views_survive_set_to_none_and_manual_grad_resets():
    """SURVEY.md hard part 3: param.grad aliases the bucket; zero_grad(set_to_none=True), `p.grad = None` and a
    replaced .grad tensor between steps must not break the averaging."""
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        torch.manual_seed(3)
        model = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.ReLU(), torch.nn.Linear(6, 1))
        ref = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.ReLU(), torch.nn.Linear(6, 1))
        ref.load_state_dict(model.state_dict())
        opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.5),
                                       named_parameters=model.named_parameters(), bucket_bytes=64)
        ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.5)
        gen = torch.Generator().manual_seed(5)
        X, Y = torch.randn(6 * n, 4, generator=gen), torch.randn(6 * n, 1, generator=gen)
        for step in range(4):
            if step == 0:
                opt.zero_grad(set_to_none=True)
            elif step == 1:
                opt.zero_grad()
                for p in model.parameters():
                    p.grad = None                                   # user code dropping the aliases
            elif step == 2:
                opt.zero_grad()
                for p in model.parameters():
                    p.grad = torch.zeros_like(p)                    # ... or replacing them by fresh tensors
            else:
                opt.zero_grad()
            ((model(X[6 * r:6 * r + 6]) - Y[6 * r:6 * r + 6]) ** 2).mean().backward()
            opt.step()
            ropt.zero_grad()
            ((ref(X) - Y) ** 2).mean().backward()
            ropt.step()
        return max((a - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
    assert HorovodRunner(np=-2).run(main) < 1e-6


def test_sparse_embedding_gradients_are_reduced_as_dense():
    """nn.Embedding(sparse=True) under DistributedOptimizer (Horovod's sparse_as_dense): the sparse gradient lands in the
    dense bucket view and is averaged like any other."""
    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        torch.manual_seed(0)
        emb = torch.nn.Embedding(20, 4, sparse=True)
        ref = torch.nn.Embedding(20, 4)
        ref.load_state_dict(emb.state_dict())
        opt = hvd.DistributedOptimizer(torch.optim.SGD(emb.parameters(), lr=0.5), named_parameters=emb.named_parameters(),
                                       sparse_as_dense=True)
        ropt = torch.optim.SGD(ref.parameters(), lr=0.5)
        idx = torch.tensor([[1, 2, 3], [4, 5, 6], [7, 8, 9], [1, 5, 9]])
        for _ in range(3):
            opt.zero_grad()
            emb(idx[r]).pow(2).sum().backward()
            opt.step()
            ropt.zero_grad()
            (sum(ref(idx[q]).pow(2).sum() for q in range(n)) / n).backward()
            ropt.step()
        return (emb.weight - ref.weight).abs().max().item()
    assert HorovodRunner(np=-2).run(main) < 1e-6


def test_collectives_fuzz_random_shapes_dtypes_and_splits():
    """150 seeded random cases at np=3: every collective x dtype x shape (zero-sized dimensions, ragged allgather,
    uneven alltoall splits, reducescatter remainders) against a locally recomputed expectation."""
    def main(cases, seed):
        import random
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        r, n = hvd.rank(), hvd.size()
        rnd = random.Random(seed)                      # same stream on every rank
        dtypes = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int32, torch.int64, torch.uint8]
        bad = []
        def data(shape, dt, rank):
            g = torch.Generator().manual_seed((rank * 1000003 + len(shape) * 7919 + sum(int(v) * (i + 1) for i, v in enumerate(shape)) * 31 + dtypes.index(dt)) % (2 ** 31))
            t = torch.randint(0, 5, tuple(shape), generator=g)
            return t.to(dt)
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
                    ok = torch.equal(out, exp) and rs.tolist() == [splits[q][r] for q in range(n)]
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
    bad = HorovodRunner(np=-3).run(main, cases=150, seed=11)
    assert bad == [], bad[:5]


def test_shutdown_then_init_in_the_same_job():
    """Horovod supports re-init; every incarnation must bootstrap through its own store-key
    namespace (rank 0 deliberately arrives late so a stale 'shm_name' would be picked up by rank 1)."""
    def main():
        import time
        import torch
        import sparkdl.hvd as hvd
        out = []
        for gen in range(3):
            if gen and hvd.rank_hint() == 0:
                time.sleep(1.0)
            hvd.init()
            r, n = hvd.rank(), hvd.size()
            s = hvd.allreduce(torch.full((17,), float(r + gen)), op=hvd.Sum, name="reinit")
            out.append(float(s[0]))
            hvd.barrier()
            hvd.shutdown()
        return out
    res = HorovodRunner(np=-2).run(main)
    assert res == [1.0, 3.0, 5.0]


def _cache_main():
    import time
    import torch
    import sparkdl.hvd as hvd
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    names = ["w%d" % i for i in range(12)]
    res = {"ok": True, "why": []}
    for step in range(6):
        # every rank submits in its own order (negotiation must still agree on one global order)
        order = list(range(len(names)))
        if r % 2 == 1:
            order.reverse()
        if step % 2 == 1:
            order = order[3:] + order[:3]
        if r == n - 1:
            time.sleep(0.01)
        ts = {i: torch.full((5 + i,), float(r + step + i)) for i in order}
        hs = {i: hvd.allreduce_async_(ts[i], op=hvd.Sum, name=names[i]) for i in order}
        for i in order:
            hvd.synchronize(hs[i])
            exp = float(sum(q + step + i for q in range(n)))
            if ts[i][0].item() != exp:
                res["ok"] = False
                res["why"].append((step, i, ts[i][0].item(), exp))
        b = torch.full((7,), float(r))
        hvd.broadcast_(b, root_rank=step % n, name="bcast")
        if b[0].item() != float(step % n):
            res["ok"] = False
            res["why"].append(("bcast", step, b[0].item()))
    st1 = hvd.engine_stats()
    # a signature change (new shape under an old name) must renegotiate, then be cached again
    for step in range(3):
        t = torch.full((3, 4), float(r + 1))
        hvd.allreduce_(t, op=hvd.Sum, name="w0")
        if t[0, 0].item() != float(n * (n + 1) // 2):
            res["ok"] = False
            res["why"].append(("reshape", step, t[0, 0].item()))
    st2 = hvd.engine_stats()
    # ragged allgather (not cacheable) keeps working next to cached names
    g = hvd.allgather(torch.full((r + 1, 2), float(r)), name="ag")
    res["ag"] = g.shape[0]
    # join: cached tensors of the ranks that are still training are completed with zeros from the joined ranks
    if r == 0:
        last = hvd.join()
    else:
        t = torch.full((5,), 2.0)
        hvd.allreduce_(t, op=hvd.Sum, name="w0j")
        t2 = torch.full((3, 4), 1.0)
        hvd.allreduce_(t2, op=hvd.Sum, name="w0")      # cached name while rank 0 has joined
        res["join_sum"] = (t[0].item(), t2[0, 0].item())
        last = hvd.join()
    res["stats1"], res["stats2"] = st1, st2
    hvd.shutdown()
    return res if r != 0 else dict(res, rank0=True)


def test_response_cache_bitmask_fast_path():
    """repeated (name, shape, dtype) collectives are agreed through the shared-memory announce bitmask;
    out-of-order submission, a signature change, ragged allgather next to cached names, and join all still work."""
    import os
    os.environ["SPARKDL_NUM_GPUS"] = "0"
    try:
        res = HorovodRunner(np=-3).run(_cache_main)
    finally:
        del os.environ["SPARKDL_NUM_GPUS"]
    assert res["ok"], res["why"]
    s1, s2 = res["stats1"], res["stats2"]
    assert s1["cache_size"] == 13 and s1["cache_hits"] >= 5 * 13 - 13       # everything after the first step is a hit
    assert s1["negotiation_us"] > 0
    assert s2["cache_misses"] == s1["cache_misses"] + 1 and s2["cache_hits"] == s1["cache_hits"] + 2
    assert res["ag"] == 6


def test_response_cache_can_be_disabled(monkeypatch):
    monkeypatch.setenv("SPARKDL_NUM_GPUS", "0")
    monkeypatch.setenv("SPARKDL_RESPONSE_CACHE", "0")

    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        for _ in range(4):
            t = torch.ones(8)
            hvd.allreduce_(t, op=hvd.Sum, name="x")
        st = hvd.engine_stats()
        hvd.shutdown()
        return st, float(t[0])
    st, v = HorovodRunner(np=-2).run(main)
    assert st["cache_hits"] == 0 and v == 2.0
