"""T2 - numerics of every sm_100a collective kernel against a plain PyTorch fp32 reference
(SURVEY.md §4b).  ``VirtualGroup`` runs N communicators in ONE process on ONE GPU: each virtual
rank launches its kernel on its own stream and the kernels synchronise through flag words exactly
as real peers do, so the kernel logic is covered without a multi-GPU box.  The multi-process /
multi-GPU path (VMM fd exchange, multicast) is covered by tests/test_gpu_multi.py."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

DT = {torch.float32: 6, torch.bfloat16: 5, torch.float16: 4, torch.float64: 7, torch.int32: 2, torch.int64: 3}


@pytest.fixture(scope="module")
def lib():
    from spark_deep_learning_b200.ops import comm
    return comm.load()


def _group(world, max_blocks=4):
    from spark_deep_learning_b200.ops.comm import VirtualGroup
    from spark_deep_learning_b200.utils.config import get_config
    cfg = get_config(refresh=True)
    cfg.kernel_timeout_ms = 5000
    return VirtualGroup(world, cfg, max_blocks=max_blocks)


def _fill(bufs, dtype, n, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    srcs = []
    for r, b in enumerate(bufs):
        if dtype.is_floating_point:
            x = torch.randn(n, generator=g, dtype=torch.float32)
        else:
            x = torch.randint(-50, 50, (n,), generator=g).float()
        t = b.tensor(dtype, n)
        t.copy_(x.to(dtype))
        srcs.append(t.clone())
    torch.cuda.synchronize()
    return srcs


def _launch_all(vg, fn):
    for r, c in enumerate(vg.comms):
        rc = fn(r, c, vg.streams[r].cuda_stream)
        assert rc >= 0, c._err()
    torch.cuda.synchronize()
    for c in vg.comms:
        c.check_error()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("algo", [1, 2, 4])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_allreduce_sum(lib, world, algo, dtype):
    vg = _group(world)
    n = 4096 * 3 + 8 if algo == 2 else 4096 + 8          # algo 4 = LL (flag-in-data); world 1 falls back to one-shot
    bufs = vg.alloc(n * 4)
    srcs = _fill(bufs, dtype, n)
    ref = sum(s.float() for s in srcs) * 0.5
    _launch_all(vg, lambda r, c, s: lib.sdlc_allreduce(c._h, 1, bufs[r].id, 0, n, DT[dtype], 0, 0.5, 1.0, algo, s))
    tol = 1e-5 if dtype == torch.float32 else 3e-2
    outs = [b.tensor(dtype, n).float() for b in bufs]
    for o in outs:
        assert torch.allclose(o, ref, rtol=tol, atol=tol * 4), (o - ref).abs().max()
    for o in outs[1:]:
        assert torch.equal(o, outs[0]), "results must be bit-identical on every rank"
    vg.close()


@pytest.mark.parametrize("dtype,op", [(torch.int32, 0), (torch.int64, 2), (torch.float64, 3), (torch.float32, 4)])
def test_allreduce_ops_and_types(lib, dtype, op):
    world, n = 4, 2048
    vg = _group(world)
    bufs = vg.alloc(n * 8)
    srcs = _fill(bufs, dtype, n, seed=3)
    st = torch.stack([s.double() for s in srcs])
    ref = {0: st.sum(0), 2: st.min(0).values, 3: st.max(0).values, 4: st.prod(0)}[op]
    _launch_all(vg, lambda r, c, s: lib.sdlc_allreduce(c._h, 1, bufs[r].id, 0, n, DT[dtype], op, 1.0, 1.0, 0, s))
    for b in bufs:
        assert torch.allclose(b.tensor(dtype, n).double(), ref, rtol=1e-5, atol=1e-5)
    vg.close()


def test_repeated_launches_reuse_flags(lib):
    """Epoch counters: back-to-back collectives on one channel, alternating algorithms and sizes."""
    world = 4
    vg = _group(world)
    bufs = vg.alloc(4 << 20)
    for it in range(12):
        n = [1000 * 4, 70000 * 4, 256][it % 3] // 4 * 4
        srcs = _fill(bufs, torch.float32, n, seed=it)
        ref = sum(s for s in srcs)
        algo = 1 + it % 2
        _launch_all(vg, lambda r, c, s: lib.sdlc_allreduce(c._h, 1, bufs[r].id, 0, n, 6, 0, 1.0, 1.0, algo, s))
        for b in bufs:
            assert torch.allclose(b.tensor(torch.float32, n), ref, rtol=1e-5, atol=1e-5), it
    vg.close()


@pytest.mark.parametrize("wire,tol", [("bf16", 2e-2), ("fp16", 2e-3), ("fp8", 0.13)])
@pytest.mark.parametrize("world", [2, 8])
def test_allreduce_wire_formats(lib, wire, tol, world):
    """K6/K7: fp32 storage, 16-bit or block-scaled-fp8 payload on the wire, fp32 accumulation."""
    from spark_deep_learning_b200.ops.comm import WIRE
    n = 128 * 37 + 40           # not a multiple of the 128-element scale block (but of 8: one 16-byte lane)
    vg = _group(world)
    grads = [torch.empty(n, device="cuda") for _ in range(world)]
    g = torch.Generator().manual_seed(5)
    for t in grads:
        t.copy_(torch.randn(n, generator=g) * torch.logspace(-3, 1, n))   # wide dynamic range
    ref = sum(t.clone() for t in grads) / world
    mag = sum(t.abs() for t in grads) / world         # error budget scales with the addends
    wb = vg.alloc(lib.sdlc_wire_bytes(n, WIRE[wire]))
    torch.cuda.synchronize()
    _launch_all(vg, lambda r, c, s: lib.sdlc_allreduce_wire(
        c._h, 1, ctypes.c_void_p(grads[r].data_ptr()), n, 6, WIRE[wire], wb[r].id, 0, 1.0, 1.0 / world, s))
    for t in grads:
        err = (t - ref).abs() / (mag + 1e-3 * mag.max())
        assert err.max() < tol * 4 and err.mean() < tol, (wire, err.max().item(), err.mean().item())
    for t in grads[1:]:
        assert torch.equal(t, grads[0])
    vg.close()


def test_pull_allgather_and_alltoall(lib):
    world = 4
    vg = _group(world)
    st = vg.alloc(1 << 16)
    rows = [3, 5, 2, 7]
    row_bytes = 36          # deliberately not a multiple of 16
    for r, b in enumerate(st):
        b.tensor(torch.uint8, rows[r] * row_bytes).copy_(
            (torch.arange(rows[r] * row_bytes) % 251 + r).to(torch.uint8))
    outs = [torch.zeros(sum(rows) * row_bytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
    torch.cuda.synchronize()
    A = ctypes.c_longlong * 8
    nb = A(*[rows[p] * row_bytes if p < world else 0 for p in range(8)])
    pre = [sum(rows[:p]) * row_bytes for p in range(world)] + [0] * (8 - world)
    _launch_all(vg, lambda r, c, s: lib.sdlc_pull(c._h, 0, st[r].id, ctypes.c_void_p(outs[r].data_ptr()),
                                                 A(*[0] * 8), nb, A(*pre), sum(rows) * row_bytes, s))
    ref = torch.cat([(torch.arange(rows[p] * row_bytes) % 251 + p).to(torch.uint8) for p in range(world)]).cuda()
    for o in outs:
        assert torch.equal(o, ref)
    vg.close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_reducescatter(lib, dtype):
    world, n = 4, 4000
    vg = _group(world)
    st = vg.alloc(n * 4)
    srcs = _fill(st, dtype, n, seed=9)
    ref = sum(s.float() for s in srcs)
    cnts = [1003, 997, 1000, 1000]        # misaligned slices exercise the scalar path
    los = [0, 1003, 2000, 3000]
    outs = [torch.zeros(cnts[r], dtype=dtype, device="cuda") for r in range(world)]
    torch.cuda.synchronize()
    _launch_all(vg, lambda r, c, s: lib.sdlc_reducescatter(
        c._h, 0, st[r].id, ctypes.c_void_p(outs[r].data_ptr()), los[r], cnts[r], DT[dtype], 0, 1.0, 1.0, 1003, s))
    tol = 1e-5 if dtype == torch.float32 else 3e-2
    for r in range(world):
        assert torch.allclose(outs[r].float(), ref[los[r]:los[r] + cnts[r]], rtol=tol, atol=tol * 4)
    vg.close()


def test_pack_unpack_roundtrip(lib):
    vg = _group(1)
    c = vg.comms[0]
    st = vg.alloc(1 << 16)[0]
    ts = [torch.randn(n, device="cuda") for n in (5, 1024, 33, 4097)]
    outs = [torch.zeros_like(t) for t in ts]
    offs, o = [], 0
    for t in ts:
        offs.append(o)
        o += (t.numel() * 4 + 15) // 16 * 16
    U, L = ctypes.c_ulonglong * len(ts), ctypes.c_longlong * len(ts)
    s = torch.cuda.current_stream().cuda_stream
    sizes = L(*[t.numel() * 4 for t in ts])
    assert lib.sdlc_pack(c._h, st.id, U(*[t.data_ptr() for t in ts]), L(*offs), sizes, len(ts), 1, s) == 0
    assert lib.sdlc_pack(c._h, st.id, U(*[t.data_ptr() for t in outs]), L(*offs), sizes, len(ts), 0, s) == 0
    torch.cuda.synchronize()
    for a, b in zip(ts, outs):
        assert torch.equal(a, b)
    vg.close()


def test_pack_more_tensors_than_one_table(lib):
    """2500 tensors in one call: more than the 1024-entry pointer table of a launch, so the batch goes out in slices."""
    vg = _group(1)
    c = vg.comms[0]
    n = 2500
    st = vg.alloc(n * 16)[0]
    src = torch.arange(n * 4, device="cuda", dtype=torch.float32).reshape(n, 4)
    dst = torch.zeros_like(src)
    U, L = ctypes.c_ulonglong * n, ctypes.c_longlong * n
    s = torch.cuda.current_stream().cuda_stream
    offs, sizes = L(*[16 * i for i in range(n)]), L(*([16] * n))
    assert lib.sdlc_pack(c._h, st.id, U(*[src[i].data_ptr() for i in range(n)]), offs, sizes, n, 1, s) == 0
    assert lib.sdlc_pack(c._h, st.id, U(*[dst[i].data_ptr() for i in range(n)]), offs, sizes, n, 0, s) == 0
    torch.cuda.synchronize()
    assert torch.equal(src, dst)
    vg.close()


def test_limits_are_reported_as_errors(lib):
    """docs/LIMITS.md: more than 8 peers, a wire-format message that is not a multiple of 8 elements."""
    from spark_deep_learning_b200.ops import comm as C
    assert lib.sdlc_max_peers() == 8
    assert not lib.sdlc_comm_create(0, 9, 0, 1000.0, 8, 1 << 20, 0)
    vg = _group(2)
    bufs = vg.alloc(1 << 12)
    for b in bufs:
        b.tensor(torch.float32, 1000).fill_(1.0)
    with pytest.raises(C.CommError, match="multiple of 8"):
        vg.comms[0].allreduce_symmetric(bufs[0], 1001, torch.float32, wire="bf16")
    vg.close()


@pytest.mark.parametrize("kind", ["sgd", "sgd_nesterov", "adam", "adamw"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_optimizer_step_matches_torch(lib, kind, dtype):
    """The fused reduce-scatter+optimizer+all-gather kernel vs torch.optim on the averaged gradient."""
    world, n, steps = 4, 5000, 3
    vg = _group(world)
    gb, pb = vg.alloc(n * 4), vg.alloc(n * 4)
    code = DT[dtype]
    p0 = torch.randn(n)
    ref_p = torch.nn.Parameter(p0.clone().cuda())
    if kind.startswith("sgd"):
        ropt = torch.optim.SGD([ref_p], lr=0.1, momentum=0.9, weight_decay=0.01, nesterov=kind.endswith("nesterov"))
    elif kind == "adam":
        ropt = torch.optim.Adam([ref_p], lr=0.01, weight_decay=0.01)
    else:
        ropt = torch.optim.AdamW([ref_p], lr=0.01, weight_decay=0.1)
    for b in pb:
        b.tensor(dtype, n).copy_(p0.to(dtype))
    ns = lib.sdlc_fused_state_elems(vg.comms[0]._h, n, code)
    s1 = [torch.zeros(ns, device="cuda") for _ in range(world)]
    s2 = [torch.zeros(ns, device="cuda") for _ in range(world)]
    master = []
    for r in range(world):
        lo = lib.sdlc_fused_slice_start(vg.comms[r]._h, n, code)
        m = torch.zeros(ns, device="cuda")
        k = max(0, min(ns, n - lo))
        m[:k] = p0[lo:lo + k].cuda()
        master.append(m)
    if kind.startswith("sgd"):
        hv = [0, 0.1, 0.01, 0.9, 0.0, 1.0 if kind.endswith("nesterov") else 0.0, 0, 0, 0]
    else:
        hv = [1 if kind == "adam" else 2, 0.01, 0.01 if kind == "adam" else 0.1, 0, 0, 0, 0.9, 0.999, 1e-8]
    hyper = [torch.tensor([float(v) for v in hv] + [0.0] * 7, device="cuda") for _ in range(world)]
    for t in range(1, steps + 1):
        grads = _fill(gb, dtype, n, seed=100 + t)
        ref_p.grad = (sum(g.float() for g in grads) / world)
        ropt.step()
        _launch_all(vg, lambda r, c, s: lib.sdlc_fused_step(
            c._h, 1, gb[r].id, pb[r].id, n, code, ctypes.c_void_p(s1[r].data_ptr()), ctypes.c_void_p(s2[r].data_ptr()),
            ctypes.c_void_p(master[r].data_ptr()) if dtype != torch.float32 else None, 1.0 / world,
            ctypes.c_void_p(hyper[r].data_ptr()), s))
        _launch_all(vg, lambda r, c, s: lib.sdlc_step_advance(c._h, 1, ctypes.c_void_p(hyper[r].data_ptr()), s))
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    got = [b.tensor(dtype, n).float() for b in pb]
    for gt in got:
        assert torch.allclose(gt, ref_p.data, rtol=tol, atol=tol), (kind, (gt - ref_p.data).abs().max().item())
    for gt in got[1:]:
        assert torch.equal(gt, got[0])
    vg.close()


def test_kernel_timeout_sets_error_instead_of_hanging(lib):
    """A missing peer must not hang the GPU: the spin-wait times out and raises on the host."""
    from spark_deep_learning_b200.ops.comm import CommError, VirtualGroup
    from spark_deep_learning_b200.utils.config import get_config
    cfg = get_config(refresh=True)
    cfg.kernel_timeout_ms = 200
    vg = VirtualGroup(2, cfg, max_blocks=2)
    bufs = vg.alloc(4096)
    c = vg.comms[0]
    assert lib.sdlc_allreduce(c._h, 3, bufs[0].id, 0, 256, 6, 0, 1.0, 1.0, 2, vg.streams[0].cuda_stream) >= 0
    torch.cuda.synchronize()          # rank 1 never launches
    with pytest.raises(CommError):
        c.check_error()
    vg.close()
