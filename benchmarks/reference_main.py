"""The training function run by the REFERENCE arm of bench.py.

Nothing of spark_deep_learning_b200 is imported here: it is what a user of the stock
``sparkdl.HorovodRunner`` would write on this box.  Horovod itself cannot be installed offline
(BASELINE.md), so the gradient exchange uses the NCCL that Horovod would call, through
``torch.nn.parallel.DistributedDataParallel`` - labelled "NCCL-via-torch stand-in for horovod+NCCL".
Stock torchvision ResNet-50, torch.optim.SGD, bf16 autocast, channels_last.
"""


def train(model_name="resnet50", batch=256, steps=20, warmup=5, e2e_steps=None, seq_len=128):
    import os
    import time

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    torch.backends.cudnn.benchmark = True

    if model_name == "resnet50":
        import torchvision
        model = torchvision.models.resnet50(weights=None).to(dev).to(memory_format=torch.channels_last)
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        host_x = torch.randn(batch, 3, 224, 224).pin_memory()
        host_y = torch.randint(0, 1000, (batch,)).pin_memory()
    else:
        from transformers import BertConfig, BertForPreTraining
        cfg = BertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                         intermediate_size=4096, vocab_size=30522, max_position_embeddings=512)
        model = BertForPreTraining(cfg).to(dev)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
        host_x = torch.randint(0, 30522, (batch, seq_len)).pin_memory()
        host_y = torch.randint(0, 30522, (batch, seq_len)).pin_memory()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank]) if world > 1 else model
    crit = torch.nn.CrossEntropyLoss()

    def fwd(x, y):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if model_name == "resnet50":
                return crit(net(x.contiguous(memory_format=torch.channels_last)), y)
            nsl = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
            return net(input_ids=x, labels=y, next_sentence_label=nsl).loss

    dx, dy = host_x.to(dev), host_y.to(dev)

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        loss = fwd(x, y)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step(dx, dy)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(dx, dy)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)

    # end to end: pinned-host inputs copied every step, loss read back every step
    k = e2e_steps or steps
    for _ in range(2):
        step(host_x.to(dev, non_blocking=True), host_y.to(dev, non_blocking=True)).item()
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(k):
        x = host_x.to(dev, non_blocking=True)
        y = host_y.to(dev, non_blocking=True)
        loss_val = step(x, y).item()
    e3.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2.elapsed_time(e3), 0.0)

    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"dev_ms": t[0].item(), "e2e_ms": t[1].item(), "steps": steps, "e2e_steps": k,
           "h2d_bytes": host_x.numel() * host_x.element_size() + host_y.numel() * host_y.element_size(),
           "d2h_bytes": 4, "e2e_wall_ms": e2e_wall_ms, "loss": loss_val, "world": world, "batch": batch}
    if world > 1:
        dist.destroy_process_group()
    return res
