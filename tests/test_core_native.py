"""Unit tests of the C++ host reduction kernels through the C ABI (no processes, no GPU)."""
import ctypes

import numpy as np
import pytest
import torch

from spark_deep_learning_b200.parallel import core


@pytest.mark.parametrize("dt,code", [(torch.float32, core.F32), (torch.float64, core.F64), (torch.int32, core.I32),
                                     (torch.int64, core.I64), (torch.float16, core.F16), (torch.bfloat16, core.BF16),
                                     (torch.uint8, core.U8)])
@pytest.mark.parametrize("op,fn", [(core.OP_SUM, torch.add), (core.OP_MIN, torch.minimum), (core.OP_MAX, torch.maximum),
                                   (core.OP_PRODUCT, torch.mul)])
def test_reduce_into_matches_torch(dt, code, op, fn):
    lib = core.load()
    g = torch.Generator().manual_seed(1)
    a = (torch.rand(257, generator=g) * 6 + 1).to(dt)
    b = (torch.rand(257, generator=g) * 6 + 1).to(dt)
    exp = fn(a.double(), b.double()) if dt.is_floating_point else fn(a, b)
    got = a.clone()
    lib.sdl_reduce_into(ctypes.c_void_p(got.data_ptr()), ctypes.c_void_p(b.data_ptr()), 257, code, op)
    if dt in (torch.float16, torch.bfloat16):
        assert torch.allclose(got.double(), exp.to(dt).double(), rtol=1e-2, atol=1e-2)
    else:
        assert torch.allclose(got.double(), exp.double().to(dt).double() if dt == torch.uint8 else exp.double())


def test_half_conversions_round_to_nearest_even():
    lib = core.load()
    x = torch.tensor([1.0, 1.0009765625, 65504.0, 6e-8, -2.5, 0.0], dtype=torch.float16)
    y = torch.zeros_like(x)
    lib.sdl_reduce_into(ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(x.data_ptr()), x.numel(), core.F16, core.OP_SUM)
    assert torch.equal(y, x)                         # 0 + x must round-trip exactly, incl. subnormals and max
    b = torch.tensor([1.0, 3.14159, -1e-3, 3e38], dtype=torch.bfloat16)
    z = torch.zeros_like(b)
    lib.sdl_reduce_into(ctypes.c_void_p(z.data_ptr()), ctypes.c_void_p(b.data_ptr()), b.numel(), core.BF16, core.OP_SUM)
    assert torch.equal(z, b)


def test_scale_buffer():
    lib = core.load()
    t = torch.arange(10, dtype=torch.float32)
    lib.sdl_scale_buffer(ctypes.c_void_p(t.data_ptr()), 10, core.F32, 0.5)
    assert torch.equal(t, torch.arange(10, dtype=torch.float32) * 0.5)
    i = torch.arange(10, dtype=torch.int64)
    lib.sdl_scale_buffer(ctypes.c_void_p(i.data_ptr()), 10, core.I64, 2.0)
    assert torch.equal(i, torch.arange(10) * 2)


def _build_stress(tmp_path, tsan):
    import os
    import subprocess
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spark_deep_learning_b200", "csrc", "core")
    exe = str(tmp_path / ("engine_stress_tsan" if tsan else "engine_stress"))
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread"] + (["-fsanitize=thread"] if tsan else []) + \
        ["-o", exe, os.path.join(src, "tests", "engine_stress.cc")] + \
        [os.path.join(src, f) for f in ("shm.cc", "cpu_collectives.cc", "engine.cc")] + ["-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return exe if r.returncode == 0 else None


def test_engine_stress_native(tmp_path):
    """3 forked ranks x 2 submitter threads, reversed submission order on odd ranks (csrc/core/tests)."""
    import subprocess
    exe = _build_stress(tmp_path, tsan=False)
    assert exe is not None
    r = subprocess.run([exe, "3", "64"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_engine_stress_thread_sanitizer(tmp_path):
    """Same binary built with -fsanitize=thread: no data-race report from the engine."""
    import subprocess
    exe = _build_stress(tmp_path, tsan=True)
    if exe is None:
        pytest.skip("libtsan not available")
    env = dict(__import__("os").environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0")
    r = subprocess.run([exe, "3", "32"], capture_output=True, text=True, timeout=280, env=env)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-3000:]


@pytest.mark.parametrize("san", ["", "-fsanitize=address,undefined", "-fsanitize=thread"])
def test_native_cpu_histogram_check_plain_and_sanitized(tmp_path, san):
    """csrc/core/tests/gbdt_hist_check.cc: the feature-parallel histogram vs a serial reference (empty input, inactive
    rows, 1-100 features, 1-64 nodes, 1 / 3 / default threads), also under ASan + UBSan and ThreadSanitizer."""
    import os
    import subprocess
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spark_deep_learning_b200", "csrc", "core")
    exe = str(tmp_path / "gbdt_hist_check")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread"] + ([san, "-fno-omit-frame-pointer"] if san else []) + \
        ["-o", exe, os.path.join(src, "tests", "gbdt_hist_check.cc"), os.path.join(src, "gbdt_cpu.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if san:
            pytest.skip("sanitizer runtime not available: %s" % r.stderr[-200:])
        raise AssertionError(r.stderr[-2000:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]
    assert "Sanitizer" not in r.stderr, r.stderr[-3000:]


def test_native_cpu_histogram_matches_numpy():
    """sdl_gbdt_hist_cpu (feature-parallel, int64) against a plain numpy scatter-add; bad node ids are rejected."""
    import numpy as np
    import torch
    from spark_deep_learning_b200.parallel import core
    lib = core.load()
    rng = np.random.RandomState(0)
    n, F, nodes = 70000, 11, 5
    bins = torch.from_numpy(rng.randint(0, 256, size=(n, F)).astype(np.uint8))
    gh = torch.from_numpy(rng.randint(-2 ** 30, 2 ** 30, size=(n, 2)).astype(np.int32))
    pos = torch.from_numpy(rng.randint(-1, nodes, size=n).astype(np.int32))
    hist = torch.zeros(nodes * F * 256 * 2, dtype=torch.int64)
    assert lib.sdl_gbdt_hist_cpu(bins.data_ptr(), gh.data_ptr(), pos.data_ptr(), hist.data_ptr(), n, F, nodes, 0) == 0
    ref = np.zeros((nodes, F, 256, 2), dtype=np.int64)
    act = pos.numpy() >= 0
    p, b, q = pos.numpy()[act], bins.numpy()[act], gh.numpy()[act].astype(np.int64)
    for f in range(F):
        np.add.at(ref[:, f, :, 0], (p, b[:, f]), q[:, 0])
        np.add.at(ref[:, f, :, 1], (p, b[:, f]), q[:, 1])
    assert np.array_equal(hist.numpy().reshape(ref.shape), ref)
    one = torch.zeros_like(hist)                                    # single-threaded path gives the same bits
    assert lib.sdl_gbdt_hist_cpu(bins.data_ptr(), gh.data_ptr(), pos.data_ptr(), one.data_ptr(), n, F, nodes, 1) == 0
    assert torch.equal(one, hist)
    pos[17] = nodes
    assert lib.sdl_gbdt_hist_cpu(bins.data_ptr(), gh.data_ptr(), pos.data_ptr(), one.data_ptr(), n, F, nodes, 0) == -1


def test_native_cpu_histogram_property():
    """Random shapes (including empty shards, one feature, every row inactive): always equal to the numpy scatter-add."""
    import numpy as np
    import torch
    from hypothesis import given, settings, strategies as st
    from spark_deep_learning_b200.parallel import core
    lib = core.load()

    @settings(max_examples=25, deadline=None)
    @given(n=st.integers(0, 3000), F=st.integers(1, 9), nodes=st.integers(1, 6), seed=st.integers(0, 10 ** 6),
           threads=st.integers(0, 4))
    def check(n, F, nodes, seed, threads):
        rng = np.random.RandomState(seed)
        bins = torch.from_numpy(rng.randint(0, 256, size=(n, F)).astype(np.uint8))
        gh = torch.from_numpy(rng.randint(-2 ** 30, 2 ** 30, size=(n, 2)).astype(np.int32))
        pos = torch.from_numpy(rng.randint(-1, nodes, size=n).astype(np.int32))
        hist = torch.zeros(nodes * F * 256 * 2, dtype=torch.int64)
        assert lib.sdl_gbdt_hist_cpu(bins.data_ptr(), gh.data_ptr(), pos.data_ptr(), hist.data_ptr(), n, F, nodes, threads) == 0
        ref = np.zeros((nodes, F, 256, 2), dtype=np.int64)
        act = pos.numpy() >= 0
        p, b, q = pos.numpy()[act], bins.numpy()[act], gh.numpy()[act].astype(np.int64)
        for f in range(F):
            np.add.at(ref[:, f, :, 0], (p, b[:, f]), q[:, 0])
            np.add.at(ref[:, f, :, 1], (p, b[:, f]), q[:, 1])
        assert np.array_equal(hist.numpy().reshape(ref.shape), ref)
    check()
