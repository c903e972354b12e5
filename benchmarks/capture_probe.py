"""Lab probe: which part of a small-BERT training step invalidates CUDA graph capture?"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparkdl.hvd as hvd
from spark_deep_learning_b200.models import bert as B

hvd.init()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
kw = dict(vocab=2000, hidden=int(os.environ.get("HID", "256")), layers=2, heads=4, ffn=1024, max_pos=128, dropout=0.0)
m = B.BertForPreTraining(fused=True, **kw).to(torch.bfloat16).to(dev)
opt = hvd.DistributedOptimizer(torch.optim.AdamW(m.parameters(), lr=1e-3), named_parameters=m.named_parameters(), fused=True)
x = torch.randint(0, 2000, (8, 128), device=dev); y = torch.randint(0, 2000, (8, 128), device=dev)
for i in range(3):
    opt.zero_grad(); l = m(x, labels=y); l.backward(); opt.step()
torch.cuda.synchronize()
print("eager ok", float(l), flush=True)


def try_capture(name, fn):
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                fn()
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        print("CAPTURE OK  ", name, flush=True)
    except Exception as e:
        print("CAPTURE FAIL", name, repr(e)[:200], flush=True)
        torch.cuda.synchronize()


from spark_deep_learning_b200.ops import gemm, xent, layernorm
h = kw["hidden"]
xx = torch.randn(8, 128, h, device=dev).bfloat16().requires_grad_(True)
L = m.layers[0]
try_capture("embeddings fwd", lambda: m.word(x) + m.pos(torch.arange(128, device=dev).unsqueeze(0)))
try_capture("qkv_proj fwd", lambda: gemm.qkv_proj(xx, L.qkv.weight, L.qkv.bias))
try_capture("layer fwd", lambda: L(xx))
try_capture("model fwd", lambda: m(x, labels=y))
def fb_layer():
    L(xx).sum().backward()
opt.zero_grad()
try_capture("layer fwd+bwd (no optimizer step)", fb_layer)
opt.synchronize() if hasattr(opt, "synchronize") else None
try:
    opt.step()
except Exception as e:
    print("step after layer bwd:", repr(e)[:120])
def fb_model():
    opt.zero_grad(); l = m(x, labels=y); l.backward(); opt.step()
try_capture("full step", fb_model)
