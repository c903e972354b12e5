"""Lab probe: a few eager steps of a small BERT through the fused path with full tracebacks."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparkdl.hvd as hvd
from spark_deep_learning_b200.models import bert as B

hvd.init()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = B.BertForPreTraining(vocab=2000, hidden=256, layers=3, heads=4, ffn=1024, max_pos=128, dropout=0.1, fused=True).to(torch.bfloat16).to(dev)
opt = hvd.DistributedOptimizer(torch.optim.AdamW(m.parameters(), lr=1e-3), named_parameters=m.named_parameters(), fused=True,
                               bucket_bytes=int(os.environ.get("PROBE_BUCKET", "400000")))
import traceback as _tb
names = {id(p): n for n, p in m.named_parameters()}
if os.environ.get("PROBE_TRACE", "1") == "1":
    for n_, p_ in m.named_parameters():
        if n_.endswith("down.weight") or n_.endswith("up.weight"):
            orig = p_._sparkdl_grad_sink
            def mk(orig, n_):
                def wrapped(p):
                    print("SINK call for", n_, flush=True)
                    return orig(p)
                return wrapped
            p_._sparkdl_grad_sink = mk(orig, n_)
            def mk2(n_):
                def posthook(p):
                    print("AUTOGRAD post-accumulate for", n_, "grad is None:", p.grad is None, flush=True)
                return posthook
            p_.register_post_accumulate_grad_hook(mk2(n_))
            def mk3(n_):
                def prehook(g):
                    print("AUTOGRAD tensor hook (grad arrives) for", n_, None if g is None else tuple(g.shape), None if g is None else float(g.float().abs().sum()), flush=True)
                return prehook
            p_.register_hook(mk3(n_))
x = torch.randint(0, 2000, (8, 128), device=dev); y = torch.randint(0, 2000, (8, 128), device=dev)
try:
    for i in range(4):
        opt.zero_grad(); l = m(x, labels=y); l.backward(); opt.step(); print("step", i, float(l), flush=True)
    from spark_deep_learning_b200.utils.graph import GraphedTrainStep
    def step(xx, yy):
        opt.zero_grad(); l = m(xx, labels=yy); l.backward(); opt.step(); return l
    g = GraphedTrainStep(step, (x, y), warmup=2)
    for i in range(3):
        print("graph step", i, float(g(x, y)), flush=True)
    print("PROBE_OK")
except Exception:
    traceback.print_exc()
    print("PROBE_FAIL")
