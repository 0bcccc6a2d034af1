"""PARSeq — the scene-text recogniser that scores generated crops (`ParseqPredictor`, reference
sgm/modules/predictors/model.py:7-57; used by test.py:58-91 and named by configs/test.yaml:31-34) — on the sm_100a kernels.

Model (src/parseq/strhub/models/parseq/system.py:35-151, modules.py:27-133; hub entry `parseq`): a ViT encoder over 4x8
patches of a 32x128 image (128 tokens, d = 384, 12 pre-LN blocks, 6 heads of 64 -> the same tcgen05 FMHA kernel as the
UNet's self-attention), one two-stream pre-LN decoder layer (12 heads of 32) queried autoregressively for up to 26
positions, one cloze refinement pass, and a 95-way head.  Every GEMM is `udt_igemm` (exact-GELU epilogue for the MLPs),
LayerNorms are `udt_layernorm`, the decoder's masked attentions `udt_mha_masked`.  torch is used for the token-embedding
gather, argmax and mask construction (a few dozen integers per image).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops, pack

SD = Dict[str, torch.Tensor]


class ParseqB200:
    """`sd`: the PARSeq checkpoint's state_dict (keys as in `parseq-bb5792a6.pt`: encoder.*, decoder.*, head.*,
    text_embed.embedding.weight, pos_queries)"""

    def __init__(self, sd: SD, device, enc_num_heads: int = 6, dec_num_heads: int = 12, max_label_length: int = 25,
                 refine_iters: int = 1, **_ignored):
        dev = torch.device(device)
        self.device = dev
        self.enc_heads, self.dec_heads = enc_num_heads, dec_num_heads
        self.max_label_length, self.refine_iters = max_label_length, refine_iters
        f = lambda k: pack.f32(sd[k]).to(dev)
        lin = lambda k: pack.pack_linear(sd[k]).to(dev)
        w = sd["encoder.patch_embed.proj.weight"]                       # [d, 3, ph, pw]
        self.d = w.shape[0]
        self.patch = (int(w.shape[2]), int(w.shape[3]))
        # patch embedding as a GEMM over (channel, py, px)-ordered patch vectors
        self.w_patch, self.b_patch = lin("encoder.patch_embed.proj.weight"), f("encoder.patch_embed.proj.bias")
        self.pos_embed = sd["encoder.pos_embed"].detach().float().reshape(-1, self.d).to(dev)     # [n, d]
        self.blocks = []
        i = 0
        while f"encoder.blocks.{i}.norm1.weight" in sd:
            p = f"encoder.blocks.{i}."
            self.blocks.append(dict(
                g1=f(p + "norm1.weight"), b1=f(p + "norm1.bias"), w_qkv=lin(p + "attn.qkv.weight"), b_qkv=f(p + "attn.qkv.bias"),
                w_o=lin(p + "attn.proj.weight"), b_o=f(p + "attn.proj.bias"), g2=f(p + "norm2.weight"), b2=f(p + "norm2.bias"),
                w_fc1=lin(p + "mlp.fc1.weight"), b_fc1=f(p + "mlp.fc1.bias"), w_fc2=lin(p + "mlp.fc2.weight"), b_fc2=f(p + "mlp.fc2.bias")))
            i += 1
        self.enc_g, self.enc_b = f("encoder.norm.weight"), f("encoder.norm.bias")
        assert "decoder.layers.1.norm_q.weight" not in sd, "PARSeq decoders deeper than one layer are not supported"
        p = "decoder.layers.0."
        d = self.d
        sw, sb = sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"]
        cw, cb = sd[p + "cross_attn.in_proj_weight"], sd[p + "cross_attn.in_proj_bias"]
        hl = lambda t: pack.pack_linear(t).to(dev)
        self.dec = dict(
            sa_wq=hl(sw[:d]), sa_bq=pack.f32(sb[:d]).to(dev), sa_wkv=hl(sw[d:]), sa_bkv=pack.f32(sb[d:]).to(dev),
            sa_wo=lin(p + "self_attn.out_proj.weight"), sa_bo=f(p + "self_attn.out_proj.bias"),
            ca_wq=hl(cw[:d]), ca_bq=pack.f32(cb[:d]).to(dev), ca_wkv=hl(cw[d:]), ca_bkv=pack.f32(cb[d:]).to(dev),
            ca_wo=lin(p + "cross_attn.out_proj.weight"), ca_bo=f(p + "cross_attn.out_proj.bias"),
            w1=lin(p + "linear1.weight"), b1=f(p + "linear1.bias"), w2=lin(p + "linear2.weight"), b2=f(p + "linear2.bias"),
            n1=(f(p + "norm1.weight"), f(p + "norm1.bias")), n2=(f(p + "norm2.weight"), f(p + "norm2.bias")),
            nq=(f(p + "norm_q.weight"), f(p + "norm_q.bias")), nc=(f(p + "norm_c.weight"), f(p + "norm_c.bias")))
        self.dn = (f("decoder.norm.weight"), f("decoder.norm.bias"))
        self.w_head, self.b_head = lin("head.weight"), f("head.bias")
        self.n_out = sd["head.weight"].shape[0]
        self.emb = (math.sqrt(d) * sd["text_embed.embedding.weight"].detach().float()).to(dev)     # TokenEmbedding: sqrt(d) * E
        self.pos_queries = sd["pos_queries"].detach().float().reshape(-1, d).to(dev)               # [26, d]
        ntok = sd["text_embed.embedding.weight"].shape[0]
        self.eos_id, self.bos_id, self.pad_id = 0, ntok - 2, ntok - 1

    # ------------------------------------------------------------------------------------------ encoder
    def encode(self, images: torch.Tensor) -> torch.Tensor:
        """images fp32 [B, 3, H, W] (normalised) -> memory fp16 [B*n, d] (Encoder.forward, modules.py:111-122)"""
        b, c, hh, ww = images.shape
        ph, pw = self.patch
        gh, gw = hh // ph, ww // pw
        n = gh * gw
        # patch vectors in the conv weight's (channel, py, px) order — data movement only
        x = images.to(self.device).float().reshape(b, c, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(b * n, c * ph * pw)
        x = x.half().contiguous()
        pos = self.pos_embed.repeat(b, 1).half().contiguous()
        t = ops.linear(x, self.w_patch, self.b_patch, residual=pos)
        d = self.d
        for w in self.blocks:
            qkv = ops.linear(ops.layernorm(t, w["g1"], w["b1"], 1e-6), w["w_qkv"], w["b_qkv"])
            a = ops.fmha(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], b, n, n, self.enc_heads, (d // self.enc_heads) ** -0.5)
            t = ops.linear(a, w["w_o"], w["b_o"], residual=t)
            h = ops.linear(ops.layernorm(t, w["g2"], w["b2"], 1e-6), w["w_fc1"], w["b_fc1"], act=ops.UDT_ACT_GELU)
            t = ops.linear(h, w["w_fc2"], w["b_fc2"], residual=t)
        return ops.layernorm(t, self.enc_g, self.enc_b, 1e-6)

    # ------------------------------------------------------------------------------------------ decoder
    def _stream(self, tgt, tgt_norm, kv_norm, mem_kv, b, lq, lk, n_mem, mask, kpm):
        """DecoderLayer.forward_stream (modules.py:57-75) on [B*lq, d] rows"""
        w, d, hd = self.dec, self.d, self.dec_heads
        q = ops.linear(tgt_norm, w["sa_wq"], w["sa_bq"])
        kv = ops.linear(kv_norm, w["sa_wkv"], w["sa_bkv"])
        a = ops.mha_masked(q, kv[:, :d], kv[:, d:], b, lq, lk, hd, mask=mask, kpm=kpm)
        tgt = ops.linear(a, w["sa_wo"], w["sa_bo"], residual=tgt)
        q2 = ops.linear(ops.layernorm(tgt, *w["n1"], 1e-5), w["ca_wq"], w["ca_bq"])
        a2 = ops.mha_masked(q2, mem_kv[:, :d], mem_kv[:, d:], b, lq, n_mem, hd)
        tgt = ops.linear(a2, w["ca_wo"], w["ca_bo"], residual=tgt)
        h = ops.linear(ops.layernorm(tgt, *w["n2"], 1e-5), w["w1"], w["b1"], act=ops.UDT_ACT_GELU)
        return ops.linear(h, w["w2"], w["b2"], residual=tgt)

    def _decode(self, tgt: torch.Tensor, mem_kv, b, n_mem, query: torch.Tensor, query_mask, kpm) -> torch.Tensor:
        """PARSeq.decode (system.py:83-95) + the single (last) decoder layer: only the query stream is updated
        (modules.py:77-86 with update_content=False).  tgt int64 [B, L]; query fp16 [B*lq, d] -> fp32 logits [B*lq, n_out]"""
        l = tgt.shape[1]
        lq = query.shape[0] // b
        content = self.emb[tgt]                                            # [B, L, d] fp32
        if l > 1:
            content[:, 1:] += self.pos_queries[: l - 1]
        content = content.reshape(b * l, self.d).half().contiguous()
        w = self.dec
        qn = ops.layernorm(query, *w["nq"], 1e-5)
        cn = ops.layernorm(content, *w["nc"], 1e-5)
        out = self._stream(query, qn, cn, mem_kv, b, lq, l, n_mem, query_mask, kpm)
        out = ops.layernorm(out, *self.dn, 1e-5)
        return ops.linear(out, self.w_head, self.b_head, out_fp32=True)

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """PARSeq.forward, decode_ar=True, max_length=None (system.py:97-151): fp32 logits [B, <= 26, 95]"""
        b = images.shape[0]
        dev = self.device
        num_steps = self.max_label_length + 1
        memory = self.encode(images)
        n_mem = memory.shape[0] // b
        w = self.dec
        mem_kv = ops.linear(memory, w["ca_wkv"], w["ca_bkv"])              # cross-attention K | V of the memory, once
        pos_q = self.pos_queries[:num_steps].half()                        # [26, d]
        full_mask = torch.triu(torch.full((num_steps, num_steps), float("-inf"), device=dev), 1)
        tgt_in = torch.full((b, num_steps), self.pad_id, dtype=torch.long, device=dev)
        tgt_in[:, 0] = self.bos_id
        logits = []
        for i in range(num_steps):
            j = i + 1
            query = pos_q[i: j].repeat(b, 1).contiguous()                  # one query position per image
            p_i = self._decode(tgt_in[:, :j], mem_kv, b, n_mem, query, full_mask[i: j, :j].contiguous(), None)
            logits.append(p_i.view(b, 1, self.n_out))
            if j < num_steps:
                tgt_in[:, j] = p_i.argmax(-1)
                if bool((tgt_in == self.eos_id).any(dim=-1).all()):        # every word has its EOS: stop (system.py:131-133)
                    break
        logits = torch.cat(logits, dim=1)
        if self.refine_iters:
            cloze = full_mask.clone()
            cloze[torch.triu(torch.ones(num_steps, num_steps, dtype=torch.bool, device=dev), 2)] = 0
            bos = torch.full((b, 1), self.bos_id, dtype=torch.long, device=dev)
            for _ in range(self.refine_iters):
                tgt_in = torch.cat([bos, logits[:, :-1].argmax(-1)], dim=1)
                l = tgt_in.shape[1]
                kpm = ((tgt_in == self.eos_id).int().cumsum(-1) > 0).to(torch.uint8).contiguous()
                query = pos_q.repeat(b, 1).contiguous()                    # ALL positions are queried, even after an early stop
                logits = self._decode(tgt_in, mem_kv, b, n_mem, query, cloze[:, :l].contiguous(), kpm).view(b, num_steps, self.n_out)
        return logits

    __call__ = forward
