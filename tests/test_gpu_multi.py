"""T2/T4 on >= 2 real GPUs: the kernels that only exist across processes - NVLS
(``multimem.ld_reduce`` / ``multimem.st``) allreduce and broadcast, P2P one-/two-shot with tails, wire formats, the fused
reduce-scatter + optimizer + all-gather step over multicast buffers, the fused GBDT histogram, the fused
``broadcast_optimizer_state`` resume path - and the loss-curve equivalence of the whole fused training step against
torch DDP.  Everything goes through ``HorovodRunner(np=-N)`` with real VMM / multicast handle exchange.

``SPARKDL_TEST_NP`` overrides the gang size (default: min(#GPUs, 8); the 2-GPU lab call uses 2)."""
import os

import pytest

from sparkdl import HorovodRunner

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _np():
    import torch
    n = int(os.environ.get("SPARKDL_TEST_NP", "0")) or min(torch.cuda.device_count(), 8)
    return -max(2, n)


# ---------------------------------------------------------------------------------------------
# K1 / K2 / K2-NVLS on a symmetric buffer, every float dtype, sizes with tails, bit-identity across ranks
# ---------------------------------------------------------------------------------------------
def _symmetric_allreduce_main():
    import torch
    import sparkdl.hvd as hvd
    from spark_deep_learning_b200.parallel import core
    hvd.init()
    comm = hvd.comm_backend()
    r, n = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    max_bytes = 24 << 20
    buf = comm.alloc_symmetric(max_bytes)
    info = {"transport": comm.transport, "multicast": bool(comm.multicast and buf.multicast), "bad": [], "ran": {}}
    algos = ["ll", "oneshot", "twoshot"] + (["multimem"] if info["multicast"] else [])
    sizes = [1, 7, 8, 1000, 4099, 16384, 65537, (1 << 20) + 3, (4 << 20) + 5]
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 2e-2), (torch.float16, 2e-3)):
        esz = torch.empty(0, dtype=dt).element_size()
        for numel in sizes:
            idx = torch.arange(numel, device=dev)
            exp = sum(((idx * 7 + q * 3) % 11 + 1).float() * 0.25 for q in range(n)) / n
            for algo in algos:
                if algo == "oneshot" and numel * esz > (8 << 20):
                    continue
                if algo == "ll" and numel * esz > (64 << 10):      # flag-in-data path: messages of at most 64 KB
                    continue
                view = buf.tensor(dt, numel)
                view.copy_((((idx * 7 + r * 3) % 11 + 1).float() * 0.25).to(dt))
                ev = comm.allreduce_symmetric(buf, numel, dt, op=core.OP_SUM, post=1.0 / n, channel=2, algo=algo)
                comm.current_stream_wait(ev, release=True)
                out = view.clone()
                torch.cuda.synchronize()
                comm.check_error()
                ok = bool(torch.allclose(out.float(), exp, rtol=tol, atol=tol))
                # every rank must hold the same bits (the allgather phase broadcasts the owner's result)
                sig = out.view(torch.uint8).to(torch.int64).sum().reshape(1)
                sigs = hvd.allgather(sig, name="sig.%s.%d.%s" % (dt, numel, algo))
                same = bool((sigs == sigs[0]).all().item())
                info["ran"][algo] = info["ran"].get(algo, 0) + 1
                if not (ok and same):
                    info["bad"].append((str(dt), numel, algo, ok, same))
                hvd.barrier()
    # 200 back-to-back LL launches of alternating sizes without any barrier in between (slot halves alternate by parity)
    view = buf.tensor(torch.float32, 16384)
    view.fill_(1.0)
    torch.cuda.synchronize()
    hvd.barrier()
    for i in range(200):
        ev = comm.allreduce_symmetric(buf, 16384 if i % 3 else 100, torch.float32, op=core.OP_SUM, post=1.0 / n, channel=2, algo="ll")
        if i < 199:
            comm.release_event(ev)
    comm.current_stream_wait(ev, release=True)
    torch.cuda.synchronize()
    comm.check_error()
    if not bool((view == 1.0).all().item()):
        info["bad"].append(("float32", 16384, "ll-repeat", False, False))
    info["ran"]["ll-repeat"] = 200
    # op = min / max and int dtypes go through the P2P kernels
    for op_name, op, fn in (("min", core.OP_MIN, min), ("max", core.OP_MAX, max)):
        view = buf.tensor(torch.float32, 5003)
        view.fill_(float(r + 1))
        ev = comm.allreduce_symmetric(buf, 5003, torch.float32, op=op, channel=2, algo="twoshot")
        comm.current_stream_wait(ev, release=True)
        torch.cuda.synchronize()
        if float(view[5002]) != float(fn(range(1, n + 1))):
            info["bad"].append((op_name, float(view[5002])))
        hvd.barrier()
    # wire formats: fp32 storage, bf16 / fp16 / block-scaled fp8 on the wire
    g = torch.Generator(device="cpu").manual_seed(5)
    base = torch.randn(n, 1 << 20, generator=g)
    exp = base.sum(0) / n
    for wire, tol in (("bf16", 2e-2), ("fp16", 4e-3), ("fp8", 0.2)):
        view = buf.tensor(torch.float32, 1 << 20)
        view.copy_(base[r].to(dev))
        ev = comm.allreduce_symmetric(buf, 1 << 20, torch.float32, op=core.OP_SUM, post=1.0 / n, wire=wire, channel=2)
        comm.current_stream_wait(ev, release=True)
        torch.cuda.synchronize()
        comm.check_error()
        err = (view.cpu() - exp).abs().max().item() / exp.abs().max().item()
        if not err < tol:
            info["bad"].append(("wire", wire, err))
        hvd.barrier()
    info["event_pool"] = comm.event_pool_size
    info["launches"] = comm.launches
    return info


def test_symmetric_allreduce_all_algorithms_dtypes_and_tails():
    info = HorovodRunner(np=_np()).run(_symmetric_allreduce_main)
    assert not info["bad"], info
    assert info["ran"].get("oneshot", 0) > 0 and info["ran"].get("twoshot", 0) > 0
    assert info["ran"].get("ll", 0) >= 15 and info["ran"].get("ll-repeat") == 200, info
    if info["transport"] == "vmm":
        assert info["multicast"] and info["ran"].get("multimem", 0) >= 20, info   # NVLS path really ran
    assert info["event_pool"] < 64      # events are recycled, not leaked


# ---------------------------------------------------------------------------------------------
# generic hvd collectives at full gang size (engine + executor; broadcast uses multimem.st when bound)
# ---------------------------------------------------------------------------------------------
def _hvd_collectives_main():
    import torch
    import sparkdl.hvd as hvd
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    bad = []
    for root in sorted({0, n - 1, n // 2}):
        for numel in (1, 1001, (3 << 20) + 1):
            for dt in (torch.float32, torch.bfloat16, torch.int64):
                t = torch.full((numel,), r + 1, device=dev).to(dt)
                hvd.broadcast_(t, root_rank=root, name="b.%d.%d.%s" % (root, numel, dt))
                if not bool((t == (root + 1)).all().item()):
                    bad.append(("broadcast", root, numel, str(dt)))
    g = hvd.allgather(torch.full((r + 1, 5), float(r), device=dev))
    if g.shape[0] != n * (n + 1) // 2 or g[-1, 0].item() != float(n - 1):
        bad.append(("allgather",))
    big = hvd.allgather(torch.full(((1 << 20) + r, 2), float(r), device=dev, dtype=torch.bfloat16))
    if big.shape[0] != n * (1 << 20) + n * (n - 1) // 2 or big[-1, 1].item() != float(n - 1):
        bad.append(("allgather-big",))
    send = torch.cat([torch.full((d + 1, 3), float(r), device=dev) for d in range(n)])
    recv, _ = hvd.alltoall(send, splits=[d + 1 for d in range(n)])
    if recv.shape[0] != n * (r + 1) or [recv[q * (r + 1), 0].item() for q in range(n)] != [float(q) for q in range(n)]:
        bad.append(("alltoall",))
    x = torch.arange((3 * n + 2) * 4, dtype=torch.float32, device=dev).reshape(3 * n + 2, 4)
    rs = hvd.reducescatter(x, op=hvd.Sum)
    extra = (3 * n + 2) - 3 * n
    rows = 3 + (1 if r < extra else 0)
    start = 3 * r + min(r, extra)
    if not torch.equal(rs, x[start:start + rows] * n):
        bad.append(("reducescatter",))
    # > 500 outstanding asynchronous collectives before the first synchronize (event ring wrap)
    ts = [torch.full((257,), float(r + k), device=dev) for k in range(700)]
    hs = [hvd.allreduce_async_(t, op=hvd.Sum, name="many.%d" % k) for k, t in enumerate(ts)]
    for k, h in enumerate(hs):
        hvd.synchronize(h)
    torch.cuda.synchronize()
    for k in (0, 1, 350, 699):
        if ts[k][0].item() != float(sum(q + k for q in range(n))):
            bad.append(("many-async", k, ts[k][0].item()))
    hvd.comm_backend().check_error()
    return bad, hvd.comm_backend().event_pool_size


def test_hvd_collectives_full_gang():
    bad, pool = HorovodRunner(np=_np()).run(_hvd_collectives_main)
    assert not bad, bad
    assert pool >= 1


# ---------------------------------------------------------------------------------------------
# fused reduce-scatter + optimizer + all-gather over multicast buffers, bf16 params with fp32 masters
# ---------------------------------------------------------------------------------------------
def _fused_bf16_main(opt_name):
    import torch
    import sparkdl.hvd as hvd
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)

    def make():
        return torch.nn.Sequential(torch.nn.Linear(64, 515), torch.nn.GELU(), torch.nn.Linear(515, 128),
                                   torch.nn.GELU(), torch.nn.Linear(128, 8)).to(dev)
    ref = make()
    model = make()
    model.load_state_dict(ref.state_dict())
    model.to(torch.bfloat16)
    ref.load_state_dict({k: v.float() for k, v in model.state_dict().items()})   # same bf16-rounded start
    mk = {"sgd": lambda p: torch.optim.SGD(p, lr=0.05, momentum=0.9, weight_decay=1e-3),
          "adamw": lambda p: torch.optim.AdamW(p, lr=3e-3, weight_decay=0.01)}[opt_name]
    opt = hvd.DistributedOptimizer(mk(model.parameters()), named_parameters=model.named_parameters(), fused=True,
                                   bucket_bytes=50000)
    ropt = mk(ref.parameters())
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    g = torch.Generator().manual_seed(11)
    X = torch.randn(16 * n, 64, generator=g).to(dev)
    Y = torch.randn(16 * n, 8, generator=g).to(dev)
    for _ in range(6):
        opt.zero_grad()
        ((model(X[16 * r:16 * r + 16].bfloat16()).float() - Y[16 * r:16 * r + 16]) ** 2).mean().backward()
        opt.step()
        ropt.zero_grad()
        ((ref(X) - Y) ** 2).mean().backward()
        ropt.step()
    torch.cuda.synchronize()
    hvd.comm_backend().check_error()
    diff = max((a.float() - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
    scale = max(b.abs().max().item() for b in ref.parameters())
    # parameters must be bit-identical on every rank after the all-gather phase
    flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()])
    sig = hvd.allgather(flat.view(torch.int32).to(torch.int64).sum().reshape(1))
    sym = all(b["symmetric"] for b in opt.bucket_summary())
    return diff / scale, bool((sig == sig[0]).all().item()), sym, hvd.comm_backend().multicast


@pytest.mark.parametrize("nvls_min_world", ["2", "3"])
@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_fused_step_bf16_master_weights_multicast(opt_name, nvls_min_world, monkeypatch):
    # "2": the multimem (NVLS) branches of fused_step_kernel even on a 2-GPU box; "3" (default): peer-to-peer at 2 ranks
    monkeypatch.setenv("SPARKDL_NVLS_MIN_WORLD", nvls_min_world)
    rel, same, sym, _ = HorovodRunner(np=_np()).run(_fused_bf16_main, opt_name=opt_name)
    assert same and sym
    # bf16 forward/backward vs an fp32 replica of the whole batch; Adam turns bf16 noise on tiny gradients into
    # full-size (+-lr) steps, so its tolerance is wider
    assert rel < (3e-2 if opt_name == "sgd" else 0.12), rel


# ---------------------------------------------------------------------------------------------
# checkpoint -> rank 0 loads -> hvd.broadcast_optimizer_state (fused, sharded state) -> all ranks continue
# ---------------------------------------------------------------------------------------------
def _bcast_opt_state_main():
    import copy
    import torch
    import sparkdl.hvd as hvd
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)

    def make():
        return torch.nn.Sequential(torch.nn.Linear(32, 257), torch.nn.ReLU(), torch.nn.Linear(257, 8)).to(dev)

    def mkopt(m):
        return hvd.DistributedOptimizer(torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=0.01),
                                        named_parameters=m.named_parameters(), fused=True, bucket_bytes=20000)
    g = torch.Generator().manual_seed(3)
    X = torch.randn(8 * n, 32, generator=g).to(dev)[8 * r:8 * r + 8]
    Y = torch.randn(8 * n, 8, generator=g).to(dev)[8 * r:8 * r + 8]

    def step(m, o):
        o.zero_grad()
        ((m(X) - Y) ** 2).mean().backward()
        o.step()
    a = make()
    init = copy.deepcopy(a.state_dict())
    oa = mkopt(a)
    hvd.broadcast_parameters(a.state_dict(), 0)
    for _ in range(3):
        step(a, oa)
    b = make()
    b.load_state_dict(init)
    ob = mkopt(b)
    hvd.broadcast_parameters(b.state_dict(), 0)
    for _ in range(2):
        step(b, ob)
    ckpt_model, ckpt_opt = copy.deepcopy(b.state_dict()), ob.state_dict()      # collective
    c = make()                                                                   # fresh weights on every rank
    oc = mkopt(c)
    if r == 0:                                                                   # only rank 0 "reads the checkpoint"
        c.load_state_dict(ckpt_model)
        oc.load_state_dict(ckpt_opt)
    hvd.broadcast_parameters(c.state_dict(), root_rank=0)
    hvd.broadcast_optimizer_state(oc, root_rank=0)
    step(c, oc)
    torch.cuda.synchronize()
    hvd.comm_backend().check_error()
    return max((p - q).abs().max().item() for p, q in zip(a.parameters(), c.parameters()))


def test_fused_broadcast_optimizer_state_resume():
    """3 uninterrupted steps == 2 steps + checkpoint + (rank 0 loads, broadcast_*) + 1 step."""
    diff = HorovodRunner(np=_np()).run(_bcast_opt_state_main)
    assert diff < 1e-5, diff


# ---------------------------------------------------------------------------------------------
# T4: loss-curve equivalence of the whole fused path (bf16 weights + fp32 masters in the sharded fused
# optimizer, fused BN / LayerNorm / tcgen05 GEMM kernels, CUDA graph) against torch DDP (fp32 masters, bf16 autocast,
# torch.optim) from the SAME initial weights on the SAME data
# ---------------------------------------------------------------------------------------------
def _equivalence_main(model_name, steps):
    import os
    import torch
    import torch.distributed as dist
    import sparkdl.hvd as hvd
    from spark_deep_learning_b200 import models
    from spark_deep_learning_b200.models import bert as bert_mod
    from spark_deep_learning_b200.utils.graph import GraphedTrainStep
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(4321)
    g = torch.Generator().manual_seed(99 + r)                        # every rank its own shard of the data
    if model_name == "resnet":
        ours = models.resnet50(num_classes=100, fused=True, bf16_weights=True).to(dev).to(memory_format=torch.channels_last)
        ref = models.resnet50(num_classes=100, fused=False).to(dev).to(memory_format=torch.channels_last)
        ref.load_state_dict({k: v.float() for k, v in ours.state_dict().items()})
        x = torch.randn(32, 3, 96, 96, generator=g).to(dev)
        y = torch.randint(0, 100, (32,), generator=g).to(dev)
        mk = lambda p: torch.optim.SGD(p, lr=0.02, momentum=0.9, weight_decay=1e-4)   # noqa: E731
        crit = torch.nn.CrossEntropyLoss()

        def fwd_ours(xx, yy):
            return crit(ours(xx.to(dtype=torch.bfloat16, memory_format=torch.channels_last)).float(), yy)

        def fwd_ref(xx, yy):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return crit(ddp(xx.contiguous(memory_format=torch.channels_last)).float(), yy)
    else:
        kw = dict(vocab=2000, hidden=256, layers=4, heads=4, ffn=1024, max_pos=128, dropout=0.0)
        ours = bert_mod.BertForPreTraining(fused=True, **kw).to(torch.bfloat16).to(dev)
        ref = bert_mod.BertForPreTraining(fused=False, **kw).to(dev)
        ref.load_state_dict({k: v.float() for k, v in ours.state_dict().items()})
        x = torch.randint(0, 2000, (16, 128), generator=g).to(dev)
        y = torch.randint(0, 2000, (16, 128), generator=g).to(dev)
        mk = lambda p: torch.optim.AdamW(p, lr=2e-4, weight_decay=0.01)   # noqa: E731

        def fwd_ours(xx, yy):
            return ours(xx, labels=yy)

        def fwd_ref(xx, yy):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return ddp(xx, labels=yy)
    opt = hvd.DistributedOptimizer(mk(ours.parameters()), named_parameters=ours.named_parameters(), fused=True)
    hvd.broadcast_parameters(ours.state_dict(), root_rank=0)

    def step(xx, yy):
        opt.zero_grad()
        loss = fwd_ours(xx, yy)
        loss.backward()
        opt.step()
        return loss
    # the first steps run eagerly (lazy state of the optimizer, cuDNN plans), the rest replay one captured CUDA graph
    ours_curve = [float(step(x, y)) for _ in range(3)]
    graphed = GraphedTrainStep(step, (x, y), warmup=0) if os.environ.get("SPARKDL_T4_GRAPH", "1") == "1" else step
    for _ in range(steps - 3):
        ours_curve.append(float(graphed(x, y)))
    torch.cuda.synchronize()
    hvd.comm_backend().check_error()
    # the reference arm: stock DDP over NCCL in the same processes
    dist.init_process_group("nccl", device_id=dev)
    ddp = torch.nn.parallel.DistributedDataParallel(ref, device_ids=[dev.index])
    ropt = mk(ref.parameters())
    ref_curve = []
    for _ in range(steps):
        ropt.zero_grad()
        loss = fwd_ref(x, y)
        loss.backward()
        ropt.step()
        ref_curve.append(float(loss))
    torch.cuda.synchronize()
    dist.destroy_process_group()
    # rank-average both curves so the comparison is about the same global batch
    c = hvd.allreduce(torch.tensor([ours_curve, ref_curve], dtype=torch.float64), name="curves")
    return c[0].tolist(), c[1].tolist()


@pytest.mark.parametrize("model_name,steps,tol", [("bert", 20, 0.03), ("resnet", 20, 0.06)])
def test_loss_curve_matches_ddp(model_name, steps, tol):
    ours, ref = HorovodRunner(np=-2).run(_equivalence_main, model_name=model_name, steps=steps)
    print("ours", ["%.4f" % v for v in ours])
    print("ddp ", ["%.4f" % v for v in ref])
    assert abs(ours[0] - ref[0]) / ref[0] < 0.01, (ours[0], ref[0])           # same weights, same data
    assert ref[-1] < ref[0]                                                     # the reference is learning
    worst = max(abs(a - b) / b for a, b in zip(ours, ref))
    assert worst < tol, (worst, ours, ref)


# ---------------------------------------------------------------------------------------------
# K8: distributed GBDT over the fused histogram + allreduce kernel == single-GPU model (through the estimator API)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fused", ["1", "0"])
def test_distributed_gbdt_matches_single_gpu(fused, monkeypatch):
    import numpy as np
    import pandas as pd
    from sparkdl.xgboost import XgboostRegressor
    monkeypatch.setenv("SPARKDL_GBDT_FUSED", fused)
    rng = np.random.RandomState(0)
    X = rng.randn(40000, 24).astype(np.float32)
    y = (2 * X[:, 0] + np.sin(3 * X[:, 1]) + 1.5 * (X[:, 2] > 0.3) + 0.1 * rng.randn(len(X))).astype(np.float32)
    X[rng.rand(*X.shape) < 0.05] = np.nan
    df = pd.DataFrame({"features": list(X), "label": y})
    kw = dict(n_estimators=8, max_depth=5, learning_rate=0.3, use_gpu=True)
    one = XgboostRegressor(**kw).fit(df).transform(df)["prediction"].to_numpy()
    n = -_np()
    model = XgboostRegressor(num_workers=n, **kw).fit(df)
    many = model.transform(df)["prediction"].to_numpy()
    assert model.get_booster().attributes["workers"] == n
    assert np.allclose(one, many, atol=2e-3), np.abs(one - many).max()
