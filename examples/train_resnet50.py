#!/usr/bin/env python
"""Data-parallel ResNet-50 on all local GPUs - the documented HorovodRunner workflow
(`/root/reference/README.md:16-93`): test with np=-1 first, then scale out.

    python examples/train_resnet50.py --np -8 --steps 100
"""
import argparse


def train(steps, batch, lr):
    import torch
    import sparkdl.hvd as hvd
    from sparkdl.horovod.torch import LogCallback
    from spark_deep_learning_b200 import models
    from spark_deep_learning_b200.utils.data import AsyncScalar, DevicePrefetcher

    hvd.init()
    dev = torch.device("cuda", torch.cuda.current_device())
    model = models.resnet50(fused=True, bf16_weights=True).to(dev).to(memory_format=torch.channels_last)
    opt = hvd.DistributedOptimizer(
        torch.optim.SGD(model.parameters(), lr=lr * hvd.size(), momentum=0.9, weight_decay=1e-4),
        named_parameters=model.named_parameters(), fused=True)
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    cb = LogCallback(per_batch_log=False) if hvd.rank() == 0 else None
    x = torch.randn(batch, 3, 224, 224).pin_memory()            # synthetic data; plug a real loader in here
    y = torch.randint(0, 1000, (batch,)).pin_memory()
    loss_reader, last = AsyncScalar(), None
    if cb:
        cb.on_epoch_begin(0)
    for i, (xb, yb) in enumerate(DevicePrefetcher([(x, y)] * steps, dev)):
        opt.zero_grad()
        logits = model(xb.to(dtype=torch.bfloat16, memory_format=torch.channels_last))
        loss = torch.nn.functional.cross_entropy(logits.float(), yb)
        loss.backward()
        opt.step()
        loss_reader.push(loss)                                   # async D2H; never stalls the launch queue
        if i % 10 == 9:
            last = loss_reader.drain()[-1]                       # one host sync every 10 steps
            if cb:
                cb.on_batch_end(i, {"loss": last})
    last = (loss_reader.drain() or [last])[-1]
    if cb:
        cb.on_epoch_end(0, {"loss": last})
    return last


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=-1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--lr", type=float, default=0.0125)
    a = ap.parse_args()
    from sparkdl import HorovodRunner
    print("final loss:", HorovodRunner(np=a.np).run(train, steps=a.steps, batch=a.batch, lr=a.lr))
