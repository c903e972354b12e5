#!/usr/bin/env python
"""Single-GPU probe: where does a ResNet-50 training step spend its time, and what do the fused
kernels / CUDA-graph capture buy?  Writes gpurun_out/resnet_probe.json + a kernel-time table."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def bench(step, iters=15, warm=5):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (time.perf_counter() - t0) * 1e3 / iters


def make(kind, batch):
    from spark_deep_learning_b200 import models
    dev = torch.device("cuda")
    if kind == "torchvision":
        import torchvision
        m = torchvision.models.resnet50(weights=None)
    else:
        m = models.resnet50(fused=(kind == "fused"))
    m = m.to(dev).to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(batch, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (batch,), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward()
        opt.step()
        return loss
    return m, opt, x, y, step


def main():
    batch = int(os.environ.get("BATCH", "256"))
    torch.backends.cudnn.benchmark = True
    out = {"batch": batch}
    for kind in ("torchvision", "unfused", "fused"):
        m, opt, x, y, step = make(kind, batch)
        dev_ms, wall_ms = bench(step)
        out[kind] = {"dev_ms": round(dev_ms, 2), "wall_ms": round(wall_ms, 2), "img_s": round(batch / dev_ms * 1e3, 1)}
        print(kind, out[kind], flush=True)
        if kind in ("torchvision", "fused"):
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
            tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70)
            with open(os.path.join(ROOT, "gpurun_out", "kernels_%s.txt" % kind), "w") as f:
                f.write(tab)
        # whole-step CUDA graph (capturable SGD: momentum buffers exist after warm-up)
        try:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(g):
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = torch.nn.functional.cross_entropy(m(x), y)
                loss.backward()
                opt.step()
            gd, gw = bench(g.replay)
            out[kind]["graph_dev_ms"] = round(gd, 2)
            out[kind]["graph_img_s"] = round(batch / gd * 1e3, 1)
            print(kind, "graph", gd, flush=True)
        except Exception as e:  # pylint: disable=broad-except
            out[kind]["graph_error"] = str(e)[:300]
            print("graph capture failed:", e, flush=True)
        del m, opt, step
        torch.cuda.empty_cache()
    with open(os.path.join(ROOT, "gpurun_out", "resnet_probe.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
