"""BERT-large for pre-training (MLM + NSP), BASELINE.json config 3.

Every projection (QKV, attention output, the FFN block, MLM transform and the 30k-wide vocabulary decoder) runs
forward AND backward through :mod:`spark_deep_learning_b200.ops.gemm`, i.e. the hand-written CTA-pair tcgen05 GEMM:
bias / GELU fused into the forward epilogue, GELU' fused into the dgrad epilogue, and the weight gradients accumulated
straight into the DistributedOptimizer's symmetric gradient buckets by the wgrad epilogue.  Attention itself is the library
flash kernel behind ``F.scaled_dot_product_attention``.  Parameters are bf16; the fused
DistributedOptimizer keeps fp32 master weights and Adam moments, sharded 1/N per GPU.
"""

import torch
from torch import nn
import torch.nn.functional as F

from spark_deep_learning_b200.ops import gemm
from spark_deep_learning_b200.ops.layernorm import FusedDropoutAddLayerNorm


class FusedLinear(nn.Module):
    """``y = act(x @ W^T + b)`` through the tcgen05 GEMM with the bias/GELU epilogue when ``fused`` and the input is bf16 CUDA."""
    def __init__(self, cin, cout, act=None, fused=True, std=0.02):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin).normal_(0, std))
        self.bias = nn.Parameter(torch.zeros(cout))
        self.act = act
        self.fused = fused

    def forward(self, x):
        if self.fused and x.is_cuda and x.dtype == torch.bfloat16:
            return gemm.linear_act(x, self.weight, self.bias, self.act)
        y = F.linear(x, self.weight.to(x.dtype), self.bias.to(x.dtype))
        return F.gelu(y) if self.act else y


class Layer(nn.Module):
    def __init__(self, h, heads, ffn, p, fused):
        super().__init__()
        self.heads = heads
        self.qkv = FusedLinear(h, 3 * h, fused=fused)
        self.proj = FusedLinear(h, h, fused=fused)
        self.ln1 = FusedDropoutAddLayerNorm(h, p, eps=1e-12, fused=fused)
        self.up = FusedLinear(h, ffn, act="gelu", fused=fused)
        self.down = FusedLinear(ffn, h, fused=fused)
        self.ln2 = FusedDropoutAddLayerNorm(h, p, eps=1e-12, fused=fused)
        self.p = p

    def forward(self, x):
        b, s, h = x.shape
        if self.qkv.fused and x.is_cuda and x.dtype == torch.bfloat16:
            # three views of one packed buffer forward, three separate gradient buffers backward (no stack / permute)
            q, k, v = (t.view(b, s, self.heads, h // self.heads).transpose(1, 2)
                       for t in gemm.qkv_proj(x, self.qkv.weight, self.qkv.bias))
        else:
            q, k, v = self.qkv(x).view(b, s, 3, self.heads, h // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, dropout_p=self.p if self.training else 0.0)
        a = self.proj(a.transpose(1, 2).reshape(b, s, h))
        x = self.ln1(x, a)                       # LayerNorm(x + dropout(a)) in one pass
        if self.up.fused and x.is_cuda and x.dtype == torch.bfloat16:
            # one autograd node: gelu' is applied inside the dgrad GEMM of the down projection
            f = gemm.mlp(x, self.up.weight, self.up.bias, self.down.weight, self.down.bias, "gelu")
        else:
            f = self.down(self.up(x))
        return self.ln2(x, f)


class BertForPreTraining(nn.Module):
    def __init__(self, vocab=30522, hidden=1024, layers=24, heads=16, ffn=4096, max_pos=512, dropout=0.1,
                 fused=True):
        super().__init__()
        self.vocab = vocab
        self.vocab_padded = (vocab + 63) // 64 * 64          # 16-byte rows for the TMA descriptors
        self.word = nn.Embedding(self.vocab_padded, hidden)
        self.pos = nn.Embedding(max_pos, hidden)
        self.seg = nn.Embedding(2, hidden)
        nn.init.normal_(self.word.weight, 0, 0.02)
        self.ln = nn.LayerNorm(hidden, eps=1e-12)
        self.layers = nn.ModuleList([Layer(hidden, heads, ffn, dropout, fused) for _ in range(layers)])
        self.pooler = FusedLinear(hidden, hidden, fused=False)
        self.mlm_dense = FusedLinear(hidden, hidden, act="gelu", fused=fused)
        self.mlm_ln = nn.LayerNorm(hidden, eps=1e-12)
        self.mlm_bias = nn.Parameter(torch.zeros(self.vocab_padded))
        self.nsp = nn.Linear(hidden, 2)
        self.p = dropout
        self.fused = fused

    def forward(self, input_ids, labels=None, next_sentence_label=None, token_type_ids=None):
        b, s = input_ids.shape
        pos = torch.arange(s, device=input_ids.device).unsqueeze(0)
        x = self.word(input_ids) + self.pos(pos)
        x = x + self.seg(token_type_ids if token_type_ids is not None else torch.zeros_like(input_ids))
        x = F.dropout(self.ln(x), self.p, self.training)
        for layer in self.layers:
            x = layer(x)
        hmlm = self.mlm_ln(self.mlm_dense(x))
        bias = self.mlm_bias.float().clone()
        bias[self.vocab:] = -1e4                                   # padded vocabulary entries never win
        if self.fused and hmlm.is_cuda and hmlm.dtype == torch.bfloat16:
            # tied decoder: the embedding lookup also sends gradient to this matrix, so its wgrad stays on autograd
            logits = gemm.linear_act(hmlm, self.word.weight, bias, None, direct_wgrad=False)
        else:
            logits = F.linear(hmlm, self.word.weight.to(hmlm.dtype), bias.to(hmlm.dtype))
        pooled = torch.tanh(self.pooler(x[:, 0]))
        nsp_logits = self.nsp(pooled)
        if labels is None:
            return logits[..., :self.vocab], nsp_logits
        flat, lab = logits.view(-1, self.vocab_padded), labels.reshape(-1)
        if self.fused and flat.is_cuda and flat.dtype == torch.bfloat16:
            from spark_deep_learning_b200.ops import xent
            loss = xent.cross_entropy(flat, lab)          # one pass each way over the [tokens, vocab] matrix, no fp32 copy
        else:
            loss = F.cross_entropy(flat.float(), lab)
        if next_sentence_label is None:
            next_sentence_label = torch.zeros(b, dtype=torch.long, device=input_ids.device)
        return loss + F.cross_entropy(nsp_logits.float(), next_sentence_label)


def bert_large(fused=True, dtype=torch.bfloat16, **kw):
    """BERT-large (24 layers, hidden 1024, 16 heads, FFN 4096) for MLM+NSP pre-training."""
    m = BertForPreTraining(fused=fused, **kw)
    return m.to(dtype) if dtype is not None else m


def bert_tiny(fused=True, dtype=torch.bfloat16):
    """2-layer, hidden-128 BERT for smoke tests."""
    m = BertForPreTraining(vocab=1000, hidden=128, layers=2, heads=4, ffn=512, max_pos=64, fused=fused)
    return m.to(dtype) if dtype is not None else m
