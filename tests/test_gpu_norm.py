"""Numerics of the fused BN(+add)(+ReLU) kernels vs a plain PyTorch fp32 reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, b, res, relu, eps):
    xf = x.float()
    y = torch.nn.functional.batch_norm(xf, None, None, w, b, True, 0.1, eps)
    if res is not None:
        y = y + res.float()
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("c,hw,n", [(64, 56, 8), (256, 14, 16), (2048, 7, 4), (512, 28, 3), (1024, 5, 2)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("with_res,relu", [(False, True), (True, True), (False, False)])
def test_bn_act_forward_backward(c, hw, n, dtype, with_res, relu):
    from spark_deep_learning_b200.ops import norm
    torch.manual_seed(0)
    dev = "cuda"
    x = (torch.randn(n, c, hw, hw, device=dev) * 2 + 0.5).to(dtype).contiguous(memory_format=torch.channels_last)
    res = torch.randn(n, c, hw, hw, device=dev).to(dtype).contiguous(memory_format=torch.channels_last) if with_res else None
    bn = torch.nn.BatchNorm2d(c).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    assert norm.supported(x)
    x1 = x.clone().requires_grad_(True)
    r1 = res.clone().requires_grad_(True) if with_res else None
    y = norm.batch_norm_act(x1, bn, r1, relu)
    g = torch.randn_like(y)
    y.backward(g)
    # reference in fp32 on the same (rounded) inputs
    x2 = x.float().clone().requires_grad_(True)
    r2 = res.float().clone().requires_grad_(True) if with_res else None
    w2 = bn.weight.detach().clone().requires_grad_(True)
    b2 = bn.bias.detach().clone().requires_grad_(True)
    yr = _ref(x2, w2, b2, r2, relu, bn.eps)
    yr.backward(g.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-4
    assert torch.allclose(y.float(), yr, rtol=tol, atol=tol), (y.float() - yr).abs().max()
    scale = x2.grad.abs().max().item() + 1e-6
    # a pre-activation within rounding distance of 0 may land on the other side of the ReLU in the two
    # implementations (different but equally valid fp32 summation orders): exclude those few elements
    with torch.no_grad():
        pre = torch.nn.functional.batch_norm(x.float(), None, None, bn.weight, bn.bias, True, 0.0, bn.eps)
        if with_res:
            pre = pre + res.float()
        safe = (pre.abs() > 1e-3) if relu else torch.ones_like(pre, dtype=torch.bool)
    err = ((x1.grad.float() - x2.grad).abs() * safe).max().item() / scale
    assert err < (3e-2 if dtype == torch.bfloat16 else 1e-3), err
    assert torch.allclose(bn.weight.grad, w2.grad, rtol=3e-2 if dtype == torch.bfloat16 else 1e-3,
                          atol=(3e-2 if dtype == torch.bfloat16 else 1e-3) * w2.grad.abs().max().item())
    assert torch.allclose(bn.bias.grad, b2.grad, rtol=3e-2 if dtype == torch.bfloat16 else 1e-3,
                          atol=(3e-2 if dtype == torch.bfloat16 else 1e-3) * b2.grad.abs().max().item())
    if with_res:
        assert ((r1.grad.float() - r2.grad).abs() * safe).max().item() < (2e-2 if dtype == torch.bfloat16 else 1e-5)
    # running statistics follow torch semantics
    m = x.float().mean((0, 2, 3))
    assert torch.allclose(bn.running_mean, 0.1 * m, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("c,hw,n", [(256, 14, 8), (64, 28, 5), (2048, 7, 2)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("use_both", [True, False])
def test_bn_act_fork_adds_the_two_gradients_in_the_kernels(c, hw, n, dtype, use_both):
    """fork=True returns two handles on one storage; their gradients are summed inside bn_bwd_reduce / bn_bwd_apply.
    Result must equal the plain op fed with g1 + g2 (and with g1 alone when the second handle is unused)."""
    from spark_deep_learning_b200.ops import norm
    torch.manual_seed(1)
    x = (torch.randn(n, c, hw, hw, device="cuda") * 2).to(dtype).contiguous(memory_format=torch.channels_last)
    res = torch.randn(n, c, hw, hw, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    g1 = torch.randn(n, c, hw, hw, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    g2 = torch.randn(n, c, hw, hw, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    grads = []
    for fork in (True, False):
        bn = torch.nn.BatchNorm2d(c).cuda()
        xa, ra = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
        if fork:
            y1, y2 = norm.batch_norm_act(xa, bn, ra, True, fork=True)
            assert y1.data_ptr() == y2.data_ptr()
            if use_both:
                torch.autograd.backward([y1, y2], [g1, g2])
            else:
                y1.backward(g1)
        else:
            y1 = norm.batch_norm_act(xa, bn, ra, True)
            y1.backward((g1.float() + g2.float()).to(dtype) if use_both else g1)
        grads.append((y1.detach().float(), xa.grad.float(), ra.grad.float(), bn.weight.grad.clone(), bn.bias.grad.clone()))
    tol = 3e-2 if dtype == torch.bfloat16 else 1e-4      # bf16: the unfused arm rounds g1 + g2 to bf16 first
    for a, b in zip(*grads):
        assert (a - b).abs().max().item() <= tol * (b.abs().max().item() + 1e-6), (a - b).abs().max().item()


def test_bn_grad_sink_accumulates_into_the_bucket_views():
    """With a gradient sink on the BatchNorm parameters (DistributedOptimizer), bn_bwd_finalize adds dweight / dbias
    straight into ``param.grad`` and autograd sees no gradient tensor."""
    from spark_deep_learning_b200.ops import norm
    torch.manual_seed(2)
    c = 128
    x = torch.randn(4, c, 14, 14, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    g = torch.randn_like(x)
    ref_bn = torch.nn.BatchNorm2d(c).cuda()
    norm.batch_norm_act(x.clone().requires_grad_(True), ref_bn, None, True).backward(g)
    bn = torch.nn.BatchNorm2d(c).cuda()
    calls = []
    for p in (bn.weight, bn.bias):
        p.grad = torch.full_like(p, 2.0)                    # pre-existing content: the sink must add to it
        p._sparkdl_grad_sink = lambda q: calls.append(q)
    norm.batch_norm_act(x.clone().requires_grad_(True), bn, None, True).backward(g)
    assert len(calls) == 2 and calls[0] is bn.weight and calls[1] is bn.bias
    assert torch.allclose(bn.weight.grad, ref_bn.weight.grad + 2.0, rtol=1e-5, atol=1e-5)
    assert torch.allclose(bn.bias.grad, ref_bn.bias.grad + 2.0, rtol=1e-5, atol=1e-5)


def test_resnet_fused_matches_unfused():
    from spark_deep_learning_b200 import models
    torch.manual_seed(0)
    a = models.resnet18ish(num_classes=10, fused=True).cuda().to(memory_format=torch.channels_last)
    b = models.resnet18ish(num_classes=10, fused=False).cuda().to(memory_format=torch.channels_last)
    b.load_state_dict(a.state_dict())
    x = torch.randn(16, 3, 64, 64, device="cuda").contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (16,), device="cuda")
    outs = []
    for m in (a, b):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward()
        outs.append((loss.item(), m.fc.weight.grad.clone(), m.stem.conv.weight.grad.clone()))
    assert abs(outs[0][0] - outs[1][0]) < 5e-2
    for k in (1, 2):
        d = (outs[0][k] - outs[1][k]).abs().max().item() / (outs[1][k].abs().max().item() + 1e-6)
        assert d < 0.15, (k, d)


@pytest.mark.parametrize("shape", [(4, 64, 112, 112), (2, 16, 7, 9), (3, 8, 1, 5), (2, 8, 6, 10), (2, 32, 8, 12), (1, 8, 4, 4)])
def test_maxpool_3x3_s2_matches_torch(shape):
    from spark_deep_learning_b200.ops import pool
    torch.manual_seed(1)
    x = torch.randn(*shape, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    x1 = x.clone().requires_grad_(True)
    x2 = x.float().clone().requires_grad_(True)
    assert pool.supported(x1)
    y = pool.max_pool_3x3_s2(x1)
    yr = torch.nn.functional.max_pool2d(x2, 3, stride=2, padding=1)
    assert y.shape == yr.shape and torch.equal(y.float(), yr)
    g = torch.randn_like(yr)
    y.backward(g.bfloat16())
    yr.backward(g.bfloat16().float())
    # ties (equal bf16 values inside one window) may pick different-but-equal maxima: compare window sums instead
    assert torch.allclose(x1.grad.float().sum(), x2.grad.sum(), rtol=1e-2, atol=1e-1)
    same = (x1.grad.float() - x2.grad).abs() < 1e-2
    assert same.float().mean().item() > 0.999


def test_stem_space_to_depth_matches_direct_convolution():
    """7x7 / stride-2 stem as a 4x4 convolution over the space-to-depth input (ops/pool.py): same numbers, forward
    and weight gradient, for fp32 NCHW and bf16 channels-last inputs."""
    import torch.nn.functional as F
    from spark_deep_learning_b200.ops import pool
    torch.manual_seed(0)
    x = torch.randn(4, 3, 64, 96, device="cuda")
    w = (torch.randn(64, 3, 7, 7, device="cuda") * 0.05).bfloat16().requires_grad_(True)
    wr = w.detach().float().requires_grad_(True)
    ref = F.conv2d(x.bfloat16().float(), wr, stride=2, padding=3)
    g = torch.randn_like(ref)
    ref.backward(g.bfloat16().float())
    for xin in (x, x.bfloat16().contiguous(memory_format=torch.channels_last)):
        assert pool.stem_s2d_supported(xin)
        w.grad = None
        y = F.conv2d(pool.stem_s2d(xin), pool.stem_s2d_weight(w).contiguous(memory_format=torch.channels_last))
        assert y.shape == ref.shape
        err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
        assert err < 2e-2, err
        y.backward(g.bfloat16())
        gerr = (w.grad.float() - wr.grad).abs().max().item() / wr.grad.abs().max().item()
        assert gerr < 3e-2, gerr
