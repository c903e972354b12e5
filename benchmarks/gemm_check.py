"""Numerics of every mode of the tcgen05 GEMM (``ops/gemm.py``) against fp32 PyTorch references.

    python benchmarks/gemm_check.py fwd | bwd | wgrad | stats | autograd | all

Each case prints one line (``ok`` / ``BAD`` + max relative error) so that a single GPU call tells which descriptor /
epilogue combinations work; the process exits non-zero when any case fails.  Run it under ``timeout``: a wrong mbarrier
protocol hangs instead of failing.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_deep_learning_b200.ops import gemm  # noqa: E402

BAD = []


def rel(a, r):
    return (a.float() - r.float()).abs().max().item() / (r.float().abs().max().item() + 1e-6)


def report(tag, errs, tol=2e-2):
    ok = all(e == e and e < tol for e in errs)
    print("%-5s %s  err=%s" % ("ok" if ok else "BAD", tag, ["%.2e" % e for e in errs]), flush=True)
    if not ok:
        BAD.append(tag)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).bfloat16()


def case_fwd():
    for (m, n, k) in [(256, 256, 64), (256, 256, 128), (128, 128, 64), (512, 384, 320), (4096, 1024, 1024),
                      (1000, 3072, 1024), (72, 1024, 4096), (4096, 30528, 1024)]:
        for act in (None, "gelu", "gelu_tanh"):
            for bn in (128, 256):
                x, w, b = rnd(m, k, scale=0.5), rnd(n, k, scale=0.05), torch.randn(n, device="cuda")
                y, pre = gemm.gemm_bf16(x, w, b, act, save_pre_act=True, block_n=bn)
                torch.cuda.synchronize()
                ref_pre = x.float() @ w.float().t() + b
                ref = ref_pre if act is None else torch.nn.functional.gelu(
                    ref_pre, approximate="tanh" if act == "gelu_tanh" else "none")
                report("fwd m%d n%d k%d act=%s bn=%d" % (m, n, k, act, bn), [rel(y, ref), rel(pre, ref_pre)])
                if act is not None:
                    # second output = act'(z) instead of z (what the backward pass multiplies by)
                    y3, der = gemm.gemm_bf16(x, w, b, act, save_pre_act=True, block_n=bn, save_deriv=True)
                    torch.cuda.synchronize()
                    z = ref_pre.clone().requires_grad_(True)
                    torch.nn.functional.gelu(z, approximate="tanh" if act == "gelu_tanh" else "none").sum().backward()
                    report("fwd+deriv m%d n%d k%d act=%s bn=%d" % (m, n, k, act, bn), [rel(y3, ref), rel(der, z.grad)])
                if act is None and bn == 256:
                    y2, _ = gemm.gemm_bf16(x, w, None, None, block_n=0)
                    torch.cuda.synchronize()
                    report("fwd-nobias m%d n%d k%d" % (m, n, k), [rel(y2, x.float() @ w.float().t())])


def case_bwd():
    for (m, n_out, k_in) in [(256, 256, 256), (512, 320, 448), (4096, 1024, 1024), (4096, 4096, 1024), (4096, 1024, 4096),
                             (1000, 1024, 3072)]:
        dy, w = rnd(m, n_out, scale=0.5), rnd(n_out, k_in, scale=0.05)
        dx = gemm.dgrad(dy, w)
        torch.cuda.synchronize()
        report("dgrad m%d nout%d kin%d" % (m, n_out, k_in), [rel(dx, dy.float() @ w.float())])
        for act in ("gelu", "gelu_tanh"):
            pre = rnd(m, k_in, scale=1.5)
            dz = gemm.dgrad(dy, w, pre, act)
            torch.cuda.synchronize()
            z = pre.float().requires_grad_(True)
            torch.nn.functional.gelu(z, approximate="tanh" if act == "gelu_tanh" else "none").sum().backward()
            report("dgrad+dgelu(%s) m%d nout%d kin%d" % (act, m, n_out, k_in), [rel(dz, (dy.float() @ w.float()) * z.grad)])
        der = rnd(m, k_in, scale=0.5)
        dz = gemm.dgrad(dy, w, der, "deriv")
        g2 = gemm.gelu_bwd(dy[:, :min(n_out, k_in)].contiguous(), der[:, :min(n_out, k_in)].contiguous(), "deriv")
        torch.cuda.synchronize()
        report("dgrad*deriv m%d nout%d kin%d" % (m, n_out, k_in),
               [rel(dz, (dy.float() @ w.float()) * der.float()),
                rel(g2, dy[:, :min(n_out, k_in)].float() * der[:, :min(n_out, k_in)].float())])


def case_wgrad():
    for (tokens, n_out, k_in) in [(256, 256, 256), (512, 256, 128), (8192, 1024, 1024), (8192, 4096, 1024),
                                  (8192, 1024, 4096), (1000, 320, 448), (4096, 30528, 1024)]:
        dy, x = rnd(tokens, n_out, scale=0.5), rnd(tokens, k_in, scale=0.5)
        ref = dy.float().t() @ x.float()
        dw = gemm.wgrad(dy, x)
        torch.cuda.synchronize()
        report("wgrad bf16 T%d nout%d kin%d (splits %d)" % (tokens, n_out, k_in, gemm._wgrad_splits(n_out, k_in, tokens)),
               [rel(dw, ref)])
        if n_out > 8192:
            continue
        base = torch.randn(n_out, k_in, device="cuda") * ref.abs().max().item() * 0.3
        for dt in (torch.float32, torch.bfloat16):
            out = base.to(dt).clone()
            gemm.wgrad(dy, x, out=out, accumulate=True)
            torch.cuda.synchronize()
            report("wgrad accumulate %s T%d nout%d kin%d" % (str(dt)[6:], tokens, n_out, k_in), [rel(out, base.to(dt).float() + ref)])
            # no split-K: accumulate inside the GEMM epilogue
            out = base.to(dt).clone()
            gemm.gemm_ex(dy, x, n_out, k_in, tokens, a_mn=True, b_mn=True,
                         epi=gemm.EPI_ACC_F32 if dt == torch.float32 else gemm.EPI_ACC_BF16, out=out, accumulate=True, splits=1)
            torch.cuda.synchronize()
            report("wgrad epilogue-accumulate %s T%d nout%d kin%d" % (str(dt)[6:], tokens, n_out, k_in),
                   [rel(out, base.to(dt).float() + ref)])
    # a view into a larger flat buffer (gradient bucket): pitch == row length, offset start
    flat = torch.zeros(4096 * 1024 + 64, device="cuda", dtype=torch.bfloat16)
    view = flat[64:].view(4096, 1024)
    dy, x = rnd(2048, 4096, scale=0.5), rnd(2048, 1024, scale=0.5)
    gemm.wgrad(dy, x, out=view, accumulate=True)
    gemm.wgrad(dy, x, out=view, accumulate=True)
    torch.cuda.synchronize()
    report("wgrad into bucket view, twice", [rel(view, 2 * (dy.float().t() @ x.float()))])
    assert float(flat[:64].abs().sum()) == 0.0


def case_stats():
    for (m, cin, cout) in [(256, 64, 64), (3136, 64, 256), (802816 // 16, 256, 64), (6272, 1024, 2048), (1000, 128, 512)]:
        x, w = rnd(m, cin, scale=0.5), rnd(cout, cin, scale=0.1)
        y, tot = gemm.conv1x1_stats(x, w)
        torch.cuda.synchronize()
        ref = x.float() @ w.float().t()
        yf = y.float()
        report("conv1x1+stats m%d cin%d cout%d" % (m, cin, cout),
               [rel(y, ref), rel(tot[0], yf.sum(0)), rel(tot[1], (yf * yf).sum(0))])


def case_autograd():
    x = rnd(8, 128, 1024, scale=0.5).requires_grad_(True)
    w = rnd(4096, 1024, scale=0.05).requires_grad_(True)
    b = torch.randn(4096, device="cuda", requires_grad=True)
    y = gemm.linear_act(x, w, b, "gelu")
    g = torch.randn_like(y)
    y.backward(g)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    torch.nn.functional.gelu(torch.nn.functional.linear(xr, wr, br)).backward(g.float())
    report("linear_act autograd", [rel(x.grad, xr.grad), rel(w.grad, wr.grad), rel(b.grad, br.grad)], tol=3e-2)
    # MLP node
    x = rnd(16, 128, 1024, scale=0.5).requires_grad_(True)
    w1, w2 = rnd(4096, 1024, scale=0.03).requires_grad_(True), rnd(1024, 4096, scale=0.03).requires_grad_(True)
    b1, b2 = torch.randn(4096, device="cuda", requires_grad=True), torch.randn(1024, device="cuda", requires_grad=True)
    y = gemm.mlp(x, w1, b1, w2, b2, "gelu")
    g = torch.randn_like(y)
    y.backward(g)
    ps = [t.detach().float().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    yr = torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(ps[0], ps[1], ps[2])), ps[3], ps[4])
    yr.backward(g.float())
    report("mlp autograd", [rel(y, yr)] + [rel(a.grad, r.grad) for a, r in zip((x, w1, b1, w2, b2), ps)], tol=4e-2)
    # grad sink: the wgrad lands in .grad directly
    w = rnd(1024, 1024, scale=0.05).requires_grad_(True)
    w.grad = torch.zeros_like(w)
    fired = []
    w._sparkdl_grad_sink = lambda p: fired.append(p)
    x = rnd(2048, 1024, scale=0.5).requires_grad_(True)
    y = gemm.linear_act(x, w, None, None)
    g = torch.randn_like(y)
    y.backward(g)
    report("wgrad through the grad sink", [rel(w.grad, g.float().t() @ x.detach().float())])
    assert len(fired) == 1


def case_misc():
    for (m, n) in [(4096, 1024), (4096, 4096), (1000, 320), (8192, 30528), (128, 8)]:
        x = rnd(m, n, scale=0.5)
        ref = x.float().sum(0)
        report("colsum fp32 m%d n%d" % (m, n), [rel(gemm.colsum(x), ref)], tol=1e-4)
        base = torch.randn(n, device="cuda")
        for dt in (torch.float32, torch.bfloat16):
            out = base.to(dt).clone()
            gemm.colsum(x, out=out, accumulate=True)
            report("colsum accumulate %s m%d n%d" % (str(dt)[6:], m, n), [rel(out, base.to(dt).float() + ref)], tol=1e-2)
        a, b = gemm.colsum(x), gemm.colsum(x)
        assert torch.equal(a, b), "colsum is not deterministic"
    x, w = rnd(512, 256, scale=0.5), rnd(384, 256, scale=0.05)
    bb = torch.randn(384, device="cuda").bfloat16()
    y, _ = gemm.gemm_bf16(x, w, bb, "gelu")
    torch.cuda.synchronize()
    report("fwd with a bf16 bias", [rel(y, torch.nn.functional.gelu(x.float() @ w.float().t() + bb.float()))])
    # bias gradient through the sink
    w = rnd(1024, 1024, scale=0.05).requires_grad_(True)
    b = torch.zeros(1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w.grad, b.grad = torch.zeros_like(w), torch.zeros_like(b)
    fired = []
    w._sparkdl_grad_sink = lambda p: fired.append("w")
    b._sparkdl_grad_sink = lambda p: fired.append("b")
    x = rnd(2048, 1024, scale=0.5).requires_grad_(True)
    y = gemm.linear_act(x, w, b, None)
    g = torch.randn_like(y)
    y.backward(g)
    report("bias grad through the sink", [rel(b.grad, g.float().sum(0)), rel(w.grad, g.float().t() @ x.detach().float())])
    assert sorted(fired) == ["b", "w"], fired


def case_qkv():
    b, s_, h, heads = 4, 128, 1024, 16
    x = rnd(b, s_, h, scale=0.5).requires_grad_(True)
    w = rnd(3 * h, h, scale=0.03).requires_grad_(True)
    bias = (torch.randn(3 * h, device="cuda") * 0.1).bfloat16().requires_grad_(True)
    q, k, v = gemm.qkv_proj(x, w, bias)
    gq, gk, gv = rnd(b, s_, h), rnd(b, s_, h), rnd(b, s_, h)
    torch.autograd.backward([q, k, v], [gq, gk, gv])
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, bias))
    y = torch.nn.functional.linear(xr, wr, br)
    y.backward(torch.cat([gq, gk, gv], -1).float())
    report("qkv_proj forward", [rel(torch.cat([q, k, v], -1), y)])
    report("qkv_proj backward (split-K dgrad, split-M wgrad, bias)", [rel(x.grad, xr.grad), rel(w.grad, wr.grad), rel(bias.grad, br.grad)],
           tol=3e-2)
    # through attention, against the packed autograd formulation
    import torch.nn.functional as F
    x2 = x.detach().clone().requires_grad_(True)
    w2 = w.detach().clone().requires_grad_(True)
    b2 = bias.detach().clone().requires_grad_(True)
    d = h // heads
    q, k, v = (t.view(b, s_, heads, d).transpose(1, 2) for t in gemm.qkv_proj(x2, w2, b2))
    a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, s_, h)
    g = rnd(b, s_, h)
    a.backward(g)
    x3, w3, b3 = (t.detach().clone().requires_grad_(True) for t in (x, w, bias))
    qkv = F.linear(x3, w3, b3).view(b, s_, 3, heads, d).permute(2, 0, 3, 1, 4)
    a3 = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(b, s_, h)
    a3.backward(g)
    report("qkv_proj + SDPA vs packed autograd", [rel(a, a3), rel(x2.grad, x3.grad), rel(w2.grad, w3.grad), rel(b2.grad, b3.grad)],
           tol=4e-2)


CASES = {"qkv": case_qkv, "misc": case_misc, "fwd": case_fwd, "bwd": case_bwd, "wgrad": case_wgrad, "stats": case_stats, "autograd": case_autograd}

if __name__ == "__main__":
    which = sys.argv[1:] or ["all"]
    torch.manual_seed(0)
    for name, fn in CASES.items():
        if "all" in which or name in which:
            fn()
    print("BAD", BAD)
    print("GEMM_OK" if not BAD else "GEMM_FAIL")
    sys.exit(1 if BAD else 0)
