"""Lab probe: memory layouts that torch's SDPA (cuDNN / flash) returns for packed-QKV inputs and their gradients."""
import torch
import torch.nn.functional as F
b, s, heads, d = 32, 128, 16, 64
h = heads * d
qkv = torch.randn(b, s, 3 * h, device="cuda", dtype=torch.bfloat16, requires_grad=True)
q, k, v = qkv.view(b, s, 3, heads, d).permute(2, 0, 3, 1, 4)
print("q shape/stride", tuple(q.shape), q.stride())
for t, n in ((q, "dq"), (k, "dk"), (v, "dv")):
    t.register_hook(lambda g, n=n: print(n, "shape/stride/contig", tuple(g.shape), g.stride(), g.is_contiguous()))
for p in (0.0, 0.1):
    a = F.scaled_dot_product_attention(q, k, v, dropout_p=p)
    print("p=%.1f out shape/stride" % p, tuple(a.shape), a.stride(), "transpose(1,2) contiguous:", a.transpose(1, 2).is_contiguous())
    a.sum().backward(retain_graph=False) if p == 0.1 else None
# separate (unpacked) projections
q2 = torch.randn(b, s, h, device="cuda", dtype=torch.bfloat16, requires_grad=True)
k2 = torch.randn(b, s, h, device="cuda", dtype=torch.bfloat16, requires_grad=True)
v2 = torch.randn(b, s, h, device="cuda", dtype=torch.bfloat16, requires_grad=True)
qq, kk, vv = (t.view(b, s, heads, d).transpose(1, 2) for t in (q2, k2, v2))
for t, n in ((qq, "dq2"), (kk, "dk2"), (vv, "dv2")):
    t.register_hook(lambda g, n=n: print(n, "shape/stride/contig", tuple(g.shape), g.stride(), g.is_contiguous()))
a = F.scaled_dot_product_attention(qq, kk, vv, dropout_p=0.1)
print("unpacked out stride", a.stride(), "transpose(1,2) contiguous:", a.transpose(1, 2).is_contiguous())
a.sum().backward()
torch.cuda.synchronize()
