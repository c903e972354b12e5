"""Lab probe 2: capture fwd+bwd of individual fused ops (no optimizer involved)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from spark_deep_learning_b200.ops import gemm, xent, layernorm

dev = torch.device("cuda", 0)
h = 1024


def rnd(*s, scale=0.05, grad=True):
    return (torch.randn(*s, device=dev) * scale).bfloat16().requires_grad_(grad)


def try_capture(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        print("CAPTURE OK  ", name, flush=True)
    except Exception as e:
        print("CAPTURE FAIL", name, repr(e)[:90], flush=True)
        torch.cuda.synchronize()


x = rnd(8, 128, h, scale=0.5)
w, b = rnd(h, h), rnd(h)
try_capture("linear_act fwd+bwd", lambda: gemm.linear_act(x, w, b, None).sum().backward())
try_capture("linear_act(gelu) fwd+bwd", lambda: gemm.linear_act(x, w, b, "gelu").sum().backward())
w1, b1, w2, b2 = rnd(4 * h, h), rnd(4 * h), rnd(h, 4 * h), rnd(h)
try_capture("mlp fwd+bwd", lambda: gemm.mlp(x, w1, b1, w2, b2, "gelu").sum().backward())
wq, bq = rnd(3 * h, h), rnd(3 * h)
def qkv():
    q, k, v = gemm.qkv_proj(x, wq, bq)
    (q.sum() + k.sum() + v.sum()).backward()
try_capture("qkv_proj fwd+bwd", qkv)
def attn():
    q, k, v = (t.view(8, 128, 16, 64).transpose(1, 2) for t in gemm.qkv_proj(x, wq, bq))
    F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(8, 128, h).sum().backward()
try_capture("qkv_proj + sdpa fwd+bwd", attn)
def attn_plain():
    q, k, v = (t.view(8, 128, 16, 64).transpose(1, 2) for t in F.linear(x, wq, bq).chunk(3, -1))
    F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(8, 128, h).sum().backward()
try_capture("plain linear + sdpa fwd+bwd", attn_plain)
ln = layernorm.FusedDropoutAddLayerNorm(h, 0.1).to(dev).to(torch.bfloat16)
a = rnd(8, 128, h, scale=0.5)
try_capture("fused LN fwd+bwd", lambda: ln(x, a).sum().backward())
logits = rnd(1024, 2048, scale=2.0)
lab = torch.randint(0, 2048, (1024,), device=dev)
try_capture("xent fwd+bwd", lambda: xent.cross_entropy(logits, lab).backward())
emb = torch.nn.Embedding(2048, h).to(dev).to(torch.bfloat16)
ids = torch.randint(0, 2000, (8, 128), device=dev)
try_capture("embedding fwd+bwd", lambda: emb(ids).sum().backward())
