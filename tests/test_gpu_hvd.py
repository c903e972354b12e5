"""T2/T4 on real GPUs through the public API: HorovodRunner -> hvd.init (handle exchange) -> CUDA
collectives via the C++ engine + executor -> DistributedOptimizer (plain and fused).
``np=-1`` cases need one GPU; ``multigpu`` cases need two or more."""
import pytest

from sparkdl import HorovodRunner

pytestmark = pytest.mark.gpu


def _cuda_collectives_main():
    import torch
    import sparkdl.hvd as hvd
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    res = {"transport": hvd.comm_backend().transport, "multicast": hvd.comm_backend().multicast}
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 3e-2), (torch.float16, 1e-2)):
        for numel in (1, 1000, 300001, (9 << 20) + 3):
            t = (torch.arange(numel, device=dev) % 13 + r + 1).to(dt)
            exp = sum((torch.arange(numel, device=dev) % 13 + q + 1).float() for q in range(n)) / n
            out = hvd.allreduce(t, name="ar.%s.%d" % (dt, numel))
            res["allreduce_%s_%d" % (dt, numel)] = bool(torch.allclose(out.float(), exp, rtol=tol, atol=tol))
    outs = hvd.grouped_allreduce([torch.full((k * 7 + 1,), float(r), device=dev) for k in range(40)], op=hvd.Sum)
    res["grouped"] = all(o[0].item() == n * (n - 1) / 2.0 and o.shape[0] == k * 7 + 1 for k, o in enumerate(outs))
    res["min"] = hvd.allreduce(torch.full((5,), float(r + 1), device=dev), op=hvd.Min)[0].item() == 1.0
    res["int_sum"] = hvd.allreduce(torch.full((5,), r + 1, device=dev, dtype=torch.int64), op=hvd.Sum)[0].item() == n * (n + 1) // 2
    g = hvd.allgather(torch.full((r + 2, 3), float(r), device=dev))
    res["allgather"] = g.shape[0] == sum(q + 2 for q in range(n)) and g[-1, 0].item() == float(n - 1) and g[0, 0].item() == 0.0
    b = torch.full((100003,), float(r), device=dev)
    hvd.broadcast_(b, root_rank=n - 1)
    res["broadcast"] = bool((b == float(n - 1)).all().item())
    send = torch.cat([torch.full((d + 1, 2), float(r), device=dev) for d in range(n)])
    recv, rs = hvd.alltoall(send, splits=[d + 1 for d in range(n)])
    res["alltoall"] = recv.shape[0] == n * (r + 1) and [recv[q * (r + 1), 0].item() for q in range(n)] == [float(q) for q in range(n)]
    x = torch.arange((2 * n + 1) * 3, dtype=torch.float32, device=dev).reshape(2 * n + 1, 3)
    rsx = hvd.reducescatter(x, op=hvd.Sum)
    rows, start = 2 + (1 if r == 0 else 0), 2 * r + (1 if r > 0 else 0)
    res["reducescatter"] = bool(torch.equal(rsx, x[start:start + rows] * n))
    model = torch.nn.Linear(64, 64).to(dev)
    with torch.no_grad():
        model.weight.fill_(float(r))
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    res["bcast_params"] = bool((model.weight == 0).all().item())
    torch.cuda.synchronize()
    hvd.comm_backend().check_error()
    res["launches"] = hvd.comm_backend().launches > 0
    return res


def _check(res):
    bad = [k for k, v in res.items() if v is False and k not in ("multicast", "transport")]
    assert not bad, (bad, res)


def test_cuda_collectives_np1():
    _check(HorovodRunner(np=-1).run(_cuda_collectives_main))


@pytest.mark.multigpu
def test_cuda_collectives_np2():
    res = HorovodRunner(np=-2).run(_cuda_collectives_main)
    _check(res)
    assert res["transport"] in ("vmm", "ipc")


@pytest.mark.multigpu
def test_cuda_collectives_np2_ipc_p2p_only(monkeypatch):
    monkeypatch.setenv("SPARKDL_TRANSPORT", "ipc")
    res = HorovodRunner(np=-2).run(_cuda_collectives_main)
    _check(res)
    assert res["transport"] == "ipc" and res["multicast"] is False


def _optimizer_main(fused, opt_name, wire):
    import os
    import torch
    import sparkdl.hvd as hvd
    if wire:
        os.environ["SPARKDL_WIRE_DTYPE"] = wire
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)

    def make():
        return torch.nn.Sequential(torch.nn.Linear(32, 257), torch.nn.ReLU(), torch.nn.Linear(257, 64),
                                   torch.nn.ReLU(), torch.nn.Linear(64, 4)).to(dev)
    model, ref = make(), make()
    ref.load_state_dict(model.state_dict())
    mk = {"sgd": lambda p: torch.optim.SGD(p, lr=0.05, momentum=0.9, weight_decay=1e-3),
          "adamw": lambda p: torch.optim.AdamW(p, lr=1e-2, weight_decay=0.01)}[opt_name]
    opt = hvd.DistributedOptimizer(mk(model.parameters()), named_parameters=model.named_parameters(),
                                   fused=fused, bucket_bytes=20000)
    ropt = mk(ref.parameters())
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(8 * n, 32, generator=g).to(dev)
    Y = torch.randn(8 * n, 4, generator=g).to(dev)
    for _ in range(4):
        opt.zero_grad()
        ((model(X[8 * r:8 * r + 8]) - Y[8 * r:8 * r + 8]) ** 2).mean().backward()
        opt.step()
        ropt.zero_grad()
        ((ref(X) - Y) ** 2).mean().backward()
        ropt.step()
    torch.cuda.synchronize()
    hvd.comm_backend().check_error()
    diff = max((a - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
    scale = max(b.abs().max().item() for b in ref.parameters())
    return diff / scale, len(opt.bucket_summary()), hvd.comm_backend().launches


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_distributed_optimizer_np1(fused, opt_name):
    rel, nb, launches = HorovodRunner(np=-1).run(_optimizer_main, fused=fused, opt_name=opt_name, wire=None)
    assert rel < 1e-4 and nb >= 2 and launches > 0


@pytest.mark.multigpu
@pytest.mark.parametrize("fused,wire,tol", [(False, None, 1e-4), (True, None, 1e-4), (False, "bf16", 2e-2),
                                             (False, "fp8", 0.2)])
def test_distributed_optimizer_np2(fused, wire, tol):
    rel, nb, launches = HorovodRunner(np=-2).run(_optimizer_main, fused=fused, opt_name="sgd", wire=wire)
    assert rel < tol and nb >= 2 and launches > 0


def test_smoke_entry():
    import __graft_entry__ as g
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.abspath(g.__file__))
    out = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "smoke ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.multigpu
def test_autotune_sets_threshold(monkeypatch, tmp_path):
    monkeypatch.setenv("SPARKDL_AUTOTUNE", "1")
    monkeypatch.setenv("HOME", str(tmp_path))

    def main():
        import torch
        import sparkdl.hvd as hvd
        hvd.init()
        tuned = hvd.comm_backend().autotuned
        ok = hvd.allreduce(torch.ones(1000, device="cuda"), op=hvd.Sum)[0].item() == hvd.size()
        torch.cuda.synchronize()
        hvd.comm_backend().check_error()
        return tuned, ok
    tuned, ok = HorovodRunner(np=-2).run(main)
    assert ok and "oneshot_max_bytes" in tuned and 0 <= tuned["oneshot_max_bytes"] <= (1 << 20)


def _resume_main(opt_name):
    """3 uninterrupted steps == 2 steps + state_dict round trip + 1 step (fused, sharded optimizer state)."""
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
        base = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9) if opt_name == "sgd" else \
            torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=0.01)
        return hvd.DistributedOptimizer(base, named_parameters=m.named_parameters(), fused=True, bucket_bytes=20000)
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
    ckpt_model, ckpt_opt = copy.deepcopy(b.state_dict()), ob.state_dict()
    c = make()
    oc = mkopt(c)
    c.load_state_dict(ckpt_model)
    oc.load_state_dict(ckpt_opt)
    step(c, oc)
    torch.cuda.synchronize()
    hvd.comm_backend().check_error()
    diff = max((p - q).abs().max().item() for p, q in zip(a.parameters(), c.parameters()))
    return diff, "sparkdl_fused" in ckpt_opt


@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_fused_optimizer_checkpoint_resume_np1(opt_name):
    diff, has = HorovodRunner(np=-1).run(_resume_main, opt_name=opt_name)
    assert has and diff < 1e-5, diff


@pytest.mark.multigpu
def test_fused_optimizer_checkpoint_resume_np2():
    diff, has = HorovodRunner(np=-2).run(_resume_main, opt_name="adamw")
    assert has and diff < 1e-5, diff
