"""tcgen05 GEMM (forward / dgrad / wgrad / BN-stats epilogues) vs plain PyTorch fp32 references.  Every group runs in
a child process under a hard timeout: a wrong mbarrier protocol must fail the test, not hang the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("group", ["misc", "qkv", "fwd", "bwd", "wgrad", "stats", "autograd"])
def test_tcgen05_gemm_numerics(group):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "gemm_check.py"), group],
                         capture_output=True, text=True, timeout=280)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    print(tail)
    assert "GEMM_OK" in out.stdout, tail
