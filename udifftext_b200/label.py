"""LabelEncoder — the character-level text encoder that conditions the UNet's `t_attn` — on the sm_100a kernels.

Reference: sgm/modules/encoders/modules.py:1069-1085 (PositionalEncoding), :1088-1173 (LabelEncoder): characters ->
indices (0 = pad / unknown, 1..94 = string.printable[:-6]) -> nn.Embedding(95, D) + sinusoid PE -> a post-LN
nn.TransformerEncoder (ReLU FFN of width D, no padding mask).  Evaluated in eval mode (dropout off); see
DESIGN.md for why the reference's own inference run leaves this module's dropout on.
"""
from __future__ import annotations

import string
from typing import Dict, List, Sequence

import torch

from . import ops, pack

SD = Dict[str, torch.Tensor]
CHARSET = string.printable[:-6]  # encoders/modules.py:1097


def label_indices(labels: Sequence[str], max_len: int) -> torch.Tensor:
    """encoders/modules.py:1149-1158 — int32 [B, max_len]"""
    rows: List[List[int]] = []
    for lab in labels:
        assert len(lab) <= max_len, f"label '{lab}' longer than max_len={max_len}"
        rows.append([CHARSET.find(ch) + 1 for ch in lab] + [0] * (max_len - len(lab)))
    return torch.tensor(rows, dtype=torch.int32).reshape(len(rows), max_len)


class LabelEncoderB200:
    """`sd` keys relative to `conditioner.embedders.0.`"""

    def __init__(self, sd: SD, device, max_len: int = 12, emb_dim: int = 2048, n_heads: int = 8, n_trans_layers: int = 12,
                 **_ignored):
        dev = torch.device(device)
        self.device, self.max_len, self.emb_dim, self.n_heads = dev, max_len, emb_dim, n_heads
        f = lambda k: pack.f32(sd[k]).to(dev)
        lin = lambda k: pack.pack_linear(sd[k]).to(dev)
        self.emb = f("label_embedding.weight")
        self.pe = f("pos_embedding.pe").reshape(-1, emb_dim)[:max_len].contiguous()
        self.layers = []
        for i in range(n_trans_layers):
            p = f"encoder.layers.{i}."
            self.layers.append(dict(
                w_in=lin(p + "self_attn.in_proj_weight"), b_in=f(p + "self_attn.in_proj_bias"),
                w_o=lin(p + "self_attn.out_proj.weight"), b_o=f(p + "self_attn.out_proj.bias"),
                w1=lin(p + "linear1.weight"), b1=f(p + "linear1.bias"), w2=lin(p + "linear2.weight"), b2=f(p + "linear2.bias"),
                g1=f(p + "norm1.weight"), n1=f(p + "norm1.bias"), g2=f(p + "norm2.weight"), n2=f(p + "norm2.bias")))

    def forward(self, labels: Sequence[str]) -> torch.Tensor:
        """list[str] -> fp32 [B, max_len, emb_dim] (encoders/modules.py:1168-1173)"""
        b = len(labels)
        idx = label_indices(labels, self.max_len).to(self.device)
        x = ops.label_embed(idx, self.emb, self.pe)
        for w in self.layers:
            qkv = ops.linear(x, w["w_in"], w["b_in"])
            a = ops.mha_small(qkv, b, self.max_len, self.n_heads)
            x = ops.layernorm(ops.linear(a, w["w_o"], w["b_o"], residual=x), w["g1"], w["n1"], 1e-5)
            h = ops.linear(x, w["w1"], w["b1"], act=ops.UDT_ACT_RELU)
            x = ops.layernorm(ops.linear(h, w["w2"], w["b2"], residual=x), w["g2"], w["n2"], 1e-5)
        return x.float().view(b, self.max_len, self.emb_dim)

    __call__ = forward
