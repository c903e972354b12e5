#!/usr/bin/env python
"""Single-GPU BERT-large step breakdown: our model (tcgen05 forward GEMMs, bf16 params) vs HF eager."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def bench(step, iters=8, warm=3):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    b, s = int(os.environ.get("BATCH", "64")), 128
    dev = torch.device("cuda")
    x = torch.randint(0, 30522, (b, s), device=dev)
    from spark_deep_learning_b200 import models
    for kind in ("ours_fused", "ours_unfused"):
        m = models.bert_large(fused=(kind == "ours_fused")).to(dev)
        opt = torch.optim.AdamW(m.parameters(), lr=1e-4, fused=True)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = m(x, labels=x)
            loss.backward()
            opt.step()
        ms = bench(step)
        print(kind, "ms/step", round(ms, 2), "samples/s", round(b / ms * 1e3, 1), flush=True)
        if kind == "ours_fused":
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                for _ in range(2):
                    step()
                torch.cuda.synchronize()
            with open(os.path.join(ROOT, "gpurun_out", "kernels_bert_ours.txt"), "w") as f:
                f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=80))
        del m, opt
        torch.cuda.empty_cache()
    from transformers import BertConfig, BertForPreTraining
    cfg = BertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                     vocab_size=30522, max_position_embeddings=512)
    m = BertForPreTraining(cfg).to(dev)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    nsl = torch.zeros(b, dtype=torch.long, device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = m(input_ids=x, labels=x, next_sentence_label=nsl).loss
        loss.backward()
        opt.step()
    ms = bench(step)
    print("hf_eager ms/step", round(ms, 2), "samples/s", round(b / ms * 1e3, 1), flush=True)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    with open(os.path.join(ROOT, "gpurun_out", "kernels_bert_hf.txt"), "w") as f:
        f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=80))


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
