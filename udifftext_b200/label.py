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


def _pair(w: torch.Tensor):
    """fp32 [out, in] -> (hi, lo) packed fp16 weights with hi + lo = w to 2^-22"""
    hi = pack.pack_linear(w)
    lo = pack.pack_linear(w.float() - hi[:, : w.shape[1]].float())
    return hi, lo


class LabelEncoderB200:
    """`sd` keys relative to `conditioner.embedders.0.`.

    Precision: the label embedding conditions every sampler step of a request, so its rounding error is systematic —
    measured on BASELINE configs[1], an fp16 residual stream here (t_crossattn rel-L2 1.2e-3) was the largest single
    contribution to the decoded-pixel error (profiles/parity_r02.json).  The encoder is 7 GFLOP per string, so it runs with
    an fp32 residual stream and fp16 operand PAIRS on the same tensor-core GEMM: x W^T = hi Wh^T + lo Wh^T + hi Wl^T
    (two udt_igemm launches with fp32 output: rows [hi; lo] against Wh, rows hi against Wl), summed / normalised in fp32
    by udt_rowsum_norm_split."""

    def __init__(self, sd: SD, device, max_len: int = 12, emb_dim: int = 2048, n_heads: int = 8, n_trans_layers: int = 12,
                 **_ignored):
        dev = torch.device(device)
        self.device, self.max_len, self.emb_dim, self.n_heads = dev, max_len, emb_dim, n_heads
        f = lambda k: pack.f32(sd[k]).to(dev)
        pair = lambda k: tuple(t.to(dev) for t in _pair(sd[k]))
        self._graph = None
        self.emb = f("label_embedding.weight")
        self.pe = f("pos_embedding.pe").reshape(-1, emb_dim)[:max_len].contiguous()
        self.layers = []
        for i in range(n_trans_layers):
            p = f"encoder.layers.{i}."
            self.layers.append(dict(
                w_in=pair(p + "self_attn.in_proj_weight"), b_in=f(p + "self_attn.in_proj_bias"),
                w_o=pair(p + "self_attn.out_proj.weight"), b_o=f(p + "self_attn.out_proj.bias"),
                w1=pair(p + "linear1.weight"), b1=f(p + "linear1.bias"), w2=pair(p + "linear2.weight"), b2=f(p + "linear2.bias"),
                g1=f(p + "norm1.weight"), n1=f(p + "norm1.bias"), g2=f(p + "norm2.weight"), n2=f(p + "norm2.bias")))

    def _pair_linear(self, xp: torch.Tensor, m: int, w, bias: torch.Tensor):
        """xp fp16 [2m, K] (rows hi then lo) -> the three fp32 partial products of x W^T + b, each [m, N]"""
        g1 = ops.linear(xp, w[0], out_fp32=True)                   # [hi; lo] Wh^T
        g2 = ops.linear(xp[:m], w[1], bias, out_fp32=True)         # hi Wl^T + b
        return g1[:m], g1[m:], g2

    def forward(self, labels: Sequence[str]) -> torch.Tensor:
        """list[str] -> fp32 [B, max_len, emb_dim] (encoders/modules.py:1168-1173); the 156 launches run as one CUDA graph
        per batch size (the result is copied out of the graph's static buffer)"""
        from .graphs import GraphCache
        if self._graph is None:
            self._graph = GraphCache(self.forward_idx)
        idx = label_indices(labels, self.max_len).to(self.device)
        return self._graph(idx).clone()

    def forward_idx(self, idx: torch.Tensor) -> torch.Tensor:
        """int32 [B, max_len] character indices (device) -> fp32 [B, max_len, emb_dim]"""
        b = idx.shape[0]
        m, d = b * self.max_len, self.emb_dim
        f16 = lambda rows, cols: torch.empty((rows, cols), device=self.device, dtype=torch.float16)
        f32 = lambda rows, cols: torch.empty((rows, cols), device=self.device, dtype=torch.float32)
        x32, xp = f32(m, d), f16(2 * m, d)
        ops.label_embed(idx, self.emb, self.pe, out=xp[:m], out_f32=x32, out_lo=xp[m:])
        for w in self.layers:
            qkv = f32(m, 3 * d)
            ops.rowsum_norm_split(self._pair_linear(xp, m, w["w_in"], w["b_in"]), out_f32=qkv)
            ap = f16(2 * m, d)
            ops.mha_small_f32(qkv, b, self.max_len, self.n_heads, ap[:m], ap[m:])
            x1, x1p = f32(m, d), f16(2 * m, d)
            ops.rowsum_norm_split(self._pair_linear(ap, m, w["w_o"], w["b_o"]), res=x32, gamma=w["g1"], beta=w["n1"],
                                  out_f32=x1, out_hi=x1p[:m], out_lo=x1p[m:])
            dff = w["w1"][0].shape[0]
            hp = f16(2 * m, dff)
            ops.rowsum_norm_split(self._pair_linear(x1p, m, w["w1"], w["b1"]), relu=True, out_hi=hp[:m], out_lo=hp[m:])
            x32, xp = f32(m, d), f16(2 * m, d)
            ops.rowsum_norm_split(self._pair_linear(hp, m, w["w2"], w["b2"]), res=x1, gamma=w["g2"], beta=w["n2"],
                                  out_f32=x32, out_hi=xp[:m], out_lo=xp[m:])
        return x32.view(b, self.max_len, self.emb_dim)

    __call__ = forward
