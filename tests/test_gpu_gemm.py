"""tcgen05 GEMM + fused epilogue vs a plain PyTorch fp32 reference.  The kernel is exercised in a
child process under a hard timeout: a wrong mbarrier protocol must fail the test, not hang the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from spark_deep_learning_b200.ops import gemm
torch.manual_seed(0)
bad = []
for (m, n, k) in [(128, 128, 64), (256, 256, 128), (4096, 1024, 1024), (4096, 4096, 1024), (1000, 3072, 1024),
                  (72, 1024, 4096), (4096, 30528, 1024)]:
    for act in (None, "gelu", "gelu_tanh"):
        for bn in (128, 256):
            x = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
            w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
            b = torch.randn(n, device="cuda")
            y, pre = gemm.gemm_bf16(x, w, b, act, save_pre_act=True, block_n=bn)
            torch.cuda.synchronize()
            ref_pre = x.float() @ w.float().t() + b
            ref = ref_pre if act is None else torch.nn.functional.gelu(ref_pre, approximate="tanh" if act == "gelu_tanh" else "none")
            e1 = (y.float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
            e2 = (pre.float() - ref_pre).abs().max().item() / (ref_pre.abs().max().item() + 1e-6)
            if not (e1 < 2e-2 and e2 < 2e-2):
                bad.append((m, n, k, act, bn, e1, e2))
# autograd wrapper
x = (torch.randn(8, 128, 1024, device="cuda") * 0.5).bfloat16().requires_grad_(True)
w = (torch.randn(4096, 1024, device="cuda") * 0.05).bfloat16().requires_grad_(True)
b = torch.randn(4096, device="cuda", requires_grad=True)
y = gemm.linear_act(x, w, b, "gelu")
g = torch.randn_like(y)
y.backward(g)
xr, wr, br = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True), b.detach().clone().requires_grad_(True)
yr = torch.nn.functional.gelu(torch.nn.functional.linear(xr, wr, br))
yr.backward(g.float())
for name, a, r in (("dx", x.grad, xr.grad), ("dw", w.grad, wr.grad), ("db", b.grad, br.grad)):
    e = (a.float() - r).abs().max().item() / (r.abs().max().item() + 1e-6)
    if e > 3e-2:
        bad.append((name, e))
# throughput (informational)
m, n, k = 8192, 4096, 4096
x = torch.randn(m, k, device="cuda").bfloat16(); w = torch.randn(n, k, device="cuda").bfloat16()
for bn in (128, 256):
    for _ in range(3): gemm.gemm_bf16(x, w, None, None, block_n=bn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): gemm.gemm_bf16(x, w, None, None, block_n=bn)
    e1.record(); torch.cuda.synchronize()
    print("tcgen05 gemm bn=%%d: %%.1f TFLOP/s" %% (bn, 2.0 * m * n * k * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e12))
for _ in range(3): x @ w.t()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x @ w.t()
e1.record(); torch.cuda.synchronize()
print("cublas: %%.1f TFLOP/s" %% (2.0 * m * n * k * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e12))
print("BAD", bad)
print("GEMM_OK" if not bad else "GEMM_FAIL")
''' % ROOT


def test_tcgen05_gemm_numerics():
    out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=240)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert "GEMM_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
