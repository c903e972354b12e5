"""Fused LayerNorm(x + dropout(a)) vs a plain PyTorch fp32 reference (same dropout mask)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _unpack_mask(mask, shape):
    bits = ((mask.view(-1, 1).int() >> torch.arange(8, device=mask.device).view(1, 8)) & 1).bool()
    return bits.view(shape)


@pytest.mark.parametrize("h", [256, 768, 1024])
@pytest.mark.parametrize("p", [0.0, 0.1])
@pytest.mark.parametrize("with_a", [True, False])
def test_fused_dropout_add_layernorm(h, p, with_a):
    from spark_deep_learning_b200.ops import layernorm as L
    torch.manual_seed(0)
    t = 4 * 37
    x = torch.randn(4, 37, h, device="cuda").bfloat16()
    a = torch.randn(4, 37, h, device="cuda").bfloat16() if with_a else None
    mod = L.FusedDropoutAddLayerNorm(h, p=p, eps=1e-12).cuda().bfloat16().train()
    with torch.no_grad():
        mod.weight.uniform_(0.5, 1.5)
        mod.bias.uniform_(-0.5, 0.5)
    x1 = x.clone().requires_grad_(True)
    a1 = a.clone().requires_grad_(True) if with_a else None
    y = mod(x1, a1)
    z_saved, mask, _, stats = y.grad_fn.saved_tensors
    keep = _unpack_mask(mask, (4, 37, h)) if (with_a and p > 0) else torch.ones(4, 37, h, dtype=torch.bool, device="cuda")
    if with_a and p > 0:
        frac = 1.0 - keep.float().mean().item()
        assert abs(frac - p) < 0.02, frac
    x2 = x.float().clone().requires_grad_(True)
    a2 = a.float().clone().requires_grad_(True) if with_a else None
    w2 = mod.weight.detach().float().clone().requires_grad_(True)
    b2 = mod.bias.detach().float().clone().requires_grad_(True)
    z = x2 + (a2 * keep / (1 - p) if with_a else 0)
    yr = torch.nn.functional.layer_norm(z, (h,), w2, b2, 1e-12)
    assert torch.allclose(y.float(), yr, rtol=2e-2, atol=2e-2), (y.float() - yr).abs().max()
    g = torch.randn_like(yr)
    y.backward(g.bfloat16())
    yr.backward(g.bfloat16().float())
    sc = x2.grad.abs().max().item()
    assert (x1.grad.float() - x2.grad).abs().max().item() / sc < 3e-2
    if with_a:
        assert (a1.grad.float() - a2.grad).abs().max().item() / (a2.grad.abs().max().item() + 1e-6) < 3e-2
    assert torch.allclose(mod.weight.grad.float(), w2.grad, rtol=5e-2, atol=5e-2 * w2.grad.abs().max().item())
    assert torch.allclose(mod.bias.grad.float(), b2.grad, rtol=5e-2, atol=5e-2 * b2.grad.abs().max().item())


def test_dropout_masks_change_between_calls():
    from spark_deep_learning_b200.ops import layernorm as L
    mod = L.FusedDropoutAddLayerNorm(256, p=0.5).cuda().bfloat16().train()
    x = torch.zeros(8, 256, device="cuda").bfloat16()
    a = torch.ones(8, 256, device="cuda").bfloat16()
    y1 = mod(x, a)
    y2 = mod(x, a)
    m1, m2 = y1.grad_fn.saved_tensors[1], y2.grad_fn.saved_tensors[1]
    assert not torch.equal(m1, m2)
    assert abs(float(y1.detach().float().mean())) < 1.0


def test_fused_cross_entropy_matches_torch():
    from spark_deep_learning_b200.ops import xent
    torch.manual_seed(0)
    for (t, v) in [(64, 1000), (300, 30528), (7, 8)]:
        logits = (torch.randn(t, v, device="cuda") * 3).bfloat16().requires_grad_(True)
        labels = torch.randint(0, v, (t,), device="cuda")
        ref_in = logits.detach().float().requires_grad_(True)
        ref = torch.nn.functional.cross_entropy(ref_in, labels)
        ref.backward()
        loss = xent.cross_entropy(logits, labels)
        (loss * 2.0).backward()
        assert abs(float(loss) - float(ref)) < 2e-3 * abs(float(ref)), (float(loss), float(ref))
        err = (logits.grad.float() - 2.0 * ref_in.grad).abs().max().item() / (2.0 * ref_in.grad.abs().max().item())
        assert err < 1e-2, err
