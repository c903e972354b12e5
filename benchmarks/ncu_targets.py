#!/usr/bin/env python
"""Small single-GPU drivers for `ncu` captures of our top kernels (B200_PROFILING.md recipe).
usage: python benchmarks/ncu_targets.py gemm|bn|allreduce|fused_step|ln|pool|xent|stem|gbdt"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def gemm():
    """One launch each of the shapes the flagship models run (fwd+GELU, dgrad, wgrad), profiler replays do the rest."""
    from spark_deep_learning_b200.ops import gemm as g
    m, n, k = 8192, 4096, 4096
    x = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    b = torch.randn(n, device="cuda")
    dy = torch.randn(m, n, device="cuda").bfloat16()
    for _ in range(2):
        g.gemm_bf16(x, w, b, "gelu")                # [8192,4096] x [4096,4096]^T + bias + GELU
        g.dgrad(dy, w)                               # dX = dY W            (B operand MN-major)
        g.wgrad(dy, x)                               # dW = dY^T X          (both operands MN-major, split-K)
    # BERT-large FFN-up at the bench batch (tokens 4096, 1024 -> 4096) and the vocabulary projection
    x2 = torch.randn(4096, 1024, device="cuda").bfloat16()
    w2 = (torch.randn(4096, 1024, device="cuda") * 0.05).bfloat16()
    w3 = (torch.randn(30528, 1024, device="cuda") * 0.05).bfloat16()
    for _ in range(2):
        g.gemm_bf16(x2, w2, b, "gelu")
        g.gemm_bf16(x2, w3)
    torch.cuda.synchronize()


def xent():
    from spark_deep_learning_b200.ops import xent as X
    logits = torch.randn(4096, 30528, device="cuda").bfloat16().requires_grad_(True)
    labels = torch.randint(0, 30522, (4096,), device="cuda")
    for _ in range(3):
        loss = X.cross_entropy(logits, labels)
        loss.backward()
    torch.cuda.synchronize()


def stem():
    from spark_deep_learning_b200.ops import pool as P
    x = torch.randn(256, 3, 224, 224, device="cuda")
    for _ in range(3):
        P.stem_s2d(x)
    torch.cuda.synchronize()


def gbdt():
    from spark_deep_learning_b200.xgboost import gbdt as G
    n, f = 1000000, 64
    bins = torch.randint(0, 256, (n, f), device="cuda", dtype=torch.uint8)
    gh = torch.randint(-2 ** 20, 2 ** 20, (n, 2), device="cuda", dtype=torch.int32)
    pos = torch.randint(0, 32, (n,), device="cuda", dtype=torch.int32)
    d = G._Dist()
    for smem in ("1", "0"):
        os.environ["SPARKDL_GBDT_SMEM_HIST"] = smem
        b = G._HistBuilder(bins, d, use_fused=False)
        for _ in range(2):
            b.build(gh, pos, 32)
    torch.cuda.synchronize()


def bn():
    from spark_deep_learning_b200.ops import norm
    x = torch.randn(256, 256, 56, 56, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = torch.randn(256, 256, 56, 56, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    m = torch.nn.BatchNorm2d(256).cuda()
    for _ in range(5):
        y = norm.batch_norm_act(x, m, r, True)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()


def _vg():
    from spark_deep_learning_b200.ops.comm import VirtualGroup
    from spark_deep_learning_b200.utils.config import get_config
    return VirtualGroup(1, get_config(refresh=True), max_blocks=32)


def allreduce():
    vg = _vg()
    lib = vg.comms[0]._lib
    n = 64 << 20
    buf = vg.alloc(n * 4)[0]
    buf.tensor(torch.float32, n).fill_(1.0)
    torch.cuda.synchronize()
    s = vg.streams[0].cuda_stream
    for _ in range(5):
        lib.sdlc_allreduce(vg.comms[0]._h, 1, buf.id, 0, n, 6, 0, 1.0, 1.0, 2, s)
    torch.cuda.synchronize()


def fused_step():
    vg = _vg()
    c = vg.comms[0]
    lib = c._lib
    n = 32 << 20
    g, p = vg.alloc(n * 4)[0], vg.alloc(n * 4)[0]
    ns = lib.sdlc_fused_state_elems(c._h, n, 6)
    s1, s2 = torch.zeros(ns, device="cuda"), torch.zeros(ns, device="cuda")
    hyper = torch.tensor([2, 1e-3, 0.01, 0, 0, 0, 0.9, 0.999, 1e-8, 5.0] + [0.0] * 6, device="cuda")
    s = vg.streams[0].cuda_stream
    for _ in range(5):
        lib.sdlc_fused_step(c._h, 1, g.id, p.id, n, 6, ctypes.c_void_p(s1.data_ptr()), ctypes.c_void_p(s2.data_ptr()), None, 1.0, ctypes.c_void_p(hyper.data_ptr()), s)
    torch.cuda.synchronize()


def ln():
    from spark_deep_learning_b200.ops import layernorm as L
    mod = L.FusedDropoutAddLayerNorm(1024, p=0.1).cuda().bfloat16().train()
    x = torch.randn(64, 128, 1024, device="cuda").bfloat16().requires_grad_(True)
    a = torch.randn(64, 128, 1024, device="cuda").bfloat16().requires_grad_(True)
    for _ in range(4):
        y = mod(x, a)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()


def pool():
    from spark_deep_learning_b200.ops import pool as P
    x = torch.randn(256, 64, 112, 112, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    for _ in range(4):
        y = P.max_pool_3x3_s2(x)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()


if __name__ == "__main__":
    {"gemm": gemm, "bn": bn, "allreduce": allreduce, "fused_step": fused_step, "ln": ln, "pool": pool, "xent": xent,
     "stem": stem, "gbdt": gbdt}[sys.argv[1]]()
