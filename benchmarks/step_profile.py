#!/usr/bin/env python
"""Per-kernel time table of OUR training step (fused model + fused DistributedOptimizer, eager so that kernels keep
their names): ``python benchmarks/step_profile.py bert|resnet [batch]`` -> gpurun_out/kernels_<model>_ours.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "bert"
    import sparkdl.hvd as hvd
    from spark_deep_learning_b200 import models
    hvd.init()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if which == "bert":
        b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
        m = models.bert_large(fused=True).to(dev)
        base = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.01)
        x = torch.randint(0, 30522, (b, 128), device=dev)
        y = torch.randint(0, 30522, (b, 128), device=dev)
        fwd = lambda: m(x, labels=y)   # noqa: E731
    else:
        b = int(sys.argv[2]) if len(sys.argv) > 2 else 256
        m = models.resnet50(fused=True, bf16_weights=True).to(dev).to(memory_format=torch.channels_last)
        base = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        x = torch.randn(b, 3, 224, 224, device=dev).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        y = torch.randint(0, 1000, (b,), device=dev)
        crit = torch.nn.CrossEntropyLoss()
        fwd = lambda: crit(m(x).float(), y)   # noqa: E731
    opt = hvd.DistributedOptimizer(base, named_parameters=m.named_parameters(), fused=True)

    def step():
        opt.zero_grad()
        loss = fwd()
        loss.backward()
        opt.step()
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("%s eager ms/step %.2f" % (which, e0.elapsed_time(e1) / 5), flush=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "kernels_%s_ours.txt" % which)
    with open(path, "w") as f:
        f.write("3 steps, batch %d\n" % b)
        f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
    print(open(path).read()[:200])


if __name__ == "__main__":
    main()
