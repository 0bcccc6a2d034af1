"""UnifiedUNetModel executed on the sm_100a kernels (reference: sgm/modules/diffusionmodules/openaimodel.py:275-624,
sgm/modules/attention.py:265-416).

Host-side plan + weight repack; every tensor op is a C-ABI kernel launch (udifftext_b200.ops).  Activations are
NHWC fp16, the residual stream stays fp16, all accumulation and normalisation statistics are fp32/fp64.
Fusions relative to the reference graph:
  * GroupNorm+SiLU is one op and also performs the skip-connection `th.cat` (two-source read);
  * conv bias, the timestep-embedding add (`h + emb_out`), the 1x1 skip conv and the residual add are the
    epilogue / extra K segments of the 3x3 implicit GEMM;
  * to_q/to_k/to_v of self-attention are one GEMM (N = 3C); GEGLU is the epilogue of ff.net.0.proj;
  * every residual add of the transformer block is a GEMM epilogue;
  * the 22 emb_layers projections are one batched GEMM, the 32 t_attn K/V projections another (step invariant).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from . import ops, pack

SD = Dict[str, torch.Tensor]


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _Res:
    """ResBlock weights (openaimodel.py:149-268)"""

    def __init__(self, sd: SD, p: str, cin: int, cout: int, dev, emb_slices: list):
        self.cin, self.cout = cin, cout
        self.g1 = pack.f32(sd[p + "in_layers.0.weight"]).to(dev)
        self.b1 = pack.f32(sd[p + "in_layers.0.bias"]).to(dev)
        self.w1 = pack.pack_conv3x3(sd[p + "in_layers.2.weight"]).to(dev)
        self.cb1 = pack.f32(sd[p + "in_layers.2.bias"]).to(dev)
        self.g2 = pack.f32(sd[p + "out_layers.0.weight"]).to(dev)
        self.b2 = pack.f32(sd[p + "out_layers.0.bias"]).to(dev)
        cb2 = sd[p + "out_layers.3.bias"].float()
        self.has_skip = (p + "skip_connection.weight") in sd
        if self.has_skip:
            self.w2 = pack.pack_conv3x3(sd[p + "out_layers.3.weight"], [sd[p + "skip_connection.weight"]]).to(dev)
            cb2 = cb2 + sd[p + "skip_connection.bias"].float()
        else:
            assert cin == cout
            self.w2 = pack.pack_conv3x3(sd[p + "out_layers.3.weight"]).to(dev)
        self.cb2 = cb2.contiguous().to(dev)
        # emb_layers.1 is batched across all ResBlocks: remember our column slice
        self.emb_off = sum(w.shape[0] for w, _ in emb_slices)
        emb_slices.append((sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"]))


class _ST:
    """SpatialTransformer + BasicTransformerBlock weights (attention.py:265-416), depth 1, linear projections"""

    def __init__(self, sd: SD, p: str, c: int, head_dim: int, dev, kv_slices: list):
        self.c = c
        self.heads = c // head_dim
        f = lambda k: pack.f32(sd[p + k]).to(dev)
        lin = lambda k: pack.pack_linear(sd[p + k]).to(dev)
        self.gn_g, self.gn_b = f("norm.weight"), f("norm.bias")
        self.w_in, self.b_in = lin("proj_in.weight"), f("proj_in.bias")
        self.w_out, self.b_out = lin("proj_out.weight"), f("proj_out.bias")
        q = "transformer_blocks.0."
        assert (p + "transformer_blocks.1.norm1.weight") not in sd, "transformer_depth > 1 not supported"
        self.ln1_g, self.ln1_b = f(q + "norm1.weight"), f(q + "norm1.bias")
        self.lnt_g, self.lnt_b = f(q + "t_norm.weight"), f(q + "t_norm.bias")
        self.ln3_g, self.ln3_b = f(q + "norm3.weight"), f(q + "norm3.bias")
        self.w_qkv = torch.cat([pack.pack_linear(sd[p + q + f"attn1.to_{n}.weight"]) for n in "qkv"], dim=0).to(dev)
        self.w_o1, self.b_o1 = lin(q + "attn1.to_out.0.weight"), f(q + "attn1.to_out.0.bias")
        self.w_tq = lin(q + "t_attn.to_q.weight")
        self.w_to, self.b_to = lin(q + "t_attn.to_out.0.weight"), f(q + "t_attn.to_out.0.bias")
        self.kv_off = sum(w.shape[0] for w in kv_slices)
        kv_slices.append(sd[p + q + "t_attn.to_k.weight"])
        kv_slices.append(sd[p + q + "t_attn.to_v.weight"])
        wg, bg = pack.pack_geglu(sd[p + q + "ff.net.0.proj.weight"].float(), sd[p + q + "ff.net.0.proj.bias"].float())
        self.w_ff1, self.b_ff1 = wg.to(dev), bg.to(dev)
        self.w_ff2, self.b_ff2 = lin(q + "ff.net.2.weight"), f(q + "ff.net.2.bias")
        self.name = (p + q + "t_attn")
        self._uc_rowbias: Dict[int, torch.Tensor] = {}

    def uc_rowbias(self, nb: int) -> torch.Tensor:
        """fp32 [nb, C]: t_attn.to_out bias for the unconditional half (first nb/2 samples), 0 for the conditional half"""
        rb = self._uc_rowbias.get(nb)
        if rb is None:
            rb = torch.zeros((nb, self.c), device=self.b_to.device, dtype=torch.float32)
            rb[: nb // 2] = self.b_to
            self._uc_rowbias[nb] = rb
        return rb


class UNetB200:
    """Inference-only UnifiedUNetModel on B200.  `sd` keys are relative to `model.diffusion_model.`."""

    def __init__(self, sd: SD, device, in_channels: int = 9, out_channels: int = 4, model_channels: int = 320,
                 attention_resolutions: Sequence[int] = (4, 2, 1), num_res_blocks: int = 2,
                 channel_mult: Sequence[int] = (1, 2, 4, 4), num_head_channels: int = 64, transformer_depth: int = 1,
                 t_context_dim: int = 2048, **_ignored):
        assert transformer_depth == 1 and num_head_channels == 64, "only SD-2 style blocks (depth 1, head dim 64)"
        dev = torch.device(device)
        self.device = dev
        self.in_channels, self.out_channels, self.mc = in_channels, out_channels, model_channels
        self.t_context_dim = t_context_dim
        self.cin_pad = _round_up(in_channels, 8)
        self.cin_pad_store = 16  # udt_cfg_pack writes the 9-channel input as 16 fp16 channels (two 16-byte stores)
        emb_slices: list = []
        kv_slices: list = []
        hd = num_head_channels

        self.w_conv_in = pack.pack_conv3x3(sd["input_blocks.0.0.weight"].float(), cin_pad=self.cin_pad).to(dev)
        self.b_conv_in = pack.f32(sd["input_blocks.0.0.bias"]).to(dev)

        # ---- mirror the constructor loops of openaimodel.py:352-536 to get the block plan
        self.input_plan: List[list] = [[("conv_in",)]]
        chans = [model_channels]
        ch, ds = model_channels, 1
        idx = 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [("res", _Res(sd, f"input_blocks.{idx}.0.", ch, mult * model_channels, dev, emb_slices))]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(("st", _ST(sd, f"input_blocks.{idx}.1.", ch, hd, dev, kv_slices)))
                self.input_plan.append(layers)
                chans.append(ch)
                idx += 1
            if level != len(channel_mult) - 1:
                w = pack.pack_conv3x3(sd[f"input_blocks.{idx}.0.op.weight"]).to(dev)
                b = pack.f32(sd[f"input_blocks.{idx}.0.op.bias"]).to(dev)
                self.input_plan.append([("down", w, b, ch)])
                chans.append(ch)
                idx += 1
                ds *= 2
        self.middle_plan = [("res", _Res(sd, "middle_block.0.", ch, ch, dev, emb_slices)),
                            ("st", _ST(sd, "middle_block.1.", ch, hd, dev, kv_slices)),
                            ("res", _Res(sd, "middle_block.2.", ch, ch, dev, emb_slices))]
        self.output_plan: List[list] = []
        idx = 0
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [("res", _Res(sd, f"output_blocks.{idx}.0.", ch + ich, model_channels * mult, dev, emb_slices))]
                ch = model_channels * mult
                j = 1
                if ds in attention_resolutions:
                    layers.append(("st", _ST(sd, f"output_blocks.{idx}.{j}.", ch, hd, dev, kv_slices)))
                    j += 1
                if level and i == num_res_blocks:
                    w = pack.pack_conv3x3(sd[f"output_blocks.{idx}.{j}.conv.weight"]).to(dev)
                    b = pack.f32(sd[f"output_blocks.{idx}.{j}.conv.bias"]).to(dev)
                    w4 = [t.to(dev) for t in pack.pack_conv3x3_up2(sd[f"output_blocks.{idx}.{j}.conv.weight"])]
                    layers.append(("up", w, b, ch, w4))
                    ds //= 2
                self.output_plan.append(layers)
                idx += 1
        self.out_g, self.out_b = pack.f32(sd["out.0.weight"]).to(dev), pack.f32(sd["out.0.bias"]).to(dev)
        self.w_out = pack.pack_conv3x3(sd["out.2.weight"]).to(dev)
        self.b_out = pack.f32(sd["out.2.bias"]).to(dev)

        # ---- batched step-invariant / batch-invariant projections
        self.w_te0 = pack.pack_linear(sd["time_embed.0.weight"]).to(dev)
        self.b_te0 = pack.f32(sd["time_embed.0.bias"]).to(dev)
        self.w_te2 = pack.pack_linear(sd["time_embed.2.weight"]).to(dev)
        self.b_te2 = pack.f32(sd["time_embed.2.bias"]).to(dev)
        self.w_emb = torch.cat([pack.pack_linear(w) for w, _ in emb_slices], dim=0).to(dev)
        self.b_emb = torch.cat([b.float() for _, b in emb_slices], dim=0).contiguous().to(dev)
        self.w_kv = torch.cat([pack.pack_linear(w) for w in kv_slices], dim=0).to(dev)
        self.kv_width = self.w_kv.shape[0]
        self.emb_width = self.w_emb.shape[0]
        half = model_channels // 2
        self._freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half).to(dev)
        self._gn_ws: Dict[int, torch.Tensor] = {}     # one workspace per batch size, never reallocated (graph-stable)
        self._max_hw = 128 * 128
        self.st_layers: List[_ST] = [l[1] for blk in self.input_plan + [self.middle_plan] + self.output_plan
                                     for l in blk if l[0] == "st"]
        # attn_map_cache of the reference (openaimodel.py:542-550): one entry per t_attn layer
        self.attn_map_cache = [{"name": st.name, "heads": st.heads, "size": None, "attn_map": None} for st in self.st_layers]
        self.export_attn_maps = False
        self.up2_min_rows = 2048
        # With an all-zero unconditional context (`"label"` in force_uc_zero_embeddings, configs/test.yaml:15) the
        # bias-free to_k / to_v give K = V = 0 for the uc half: its t_attn output is exactly to_out.bias (SURVEY §7
        # (iii)).  The runner sets this flag per request after checking the context; the uc half then skips
        # t_norm / to_q / attention / to_out and receives the bias through the previous GEMM's per-sample bias.
        self.skip_uc_xattn = False
        # per-request folded t_attn weights {layer index: (W1 [B, npad, C], W2 [B, C, npad])} of the conditional half
        # (fold_context), or None: the runner owns the buffers and sets this before every forward
        self.xattn_fold: Optional[Dict[int, tuple]] = None

    # ------------------------------------------------------------------------------------------ pieces
    def _ws(self, nb: int) -> torch.Tensor:
        """GroupNorm partial-statistics workspace (shared by all calls of one batch size: they are stream ordered).
        Captured step graphs hold its raw pointer, so a workspace is never freed or replaced while the model lives:
        a larger batch gets its own buffer instead of growing (and thereby freeing) the one a cached graph writes to."""
        ws = self._gn_ws.get(nb)
        if ws is None:
            hw = self._max_hw
            need = max(ops.groupnorm_ws_bytes(nb, hw, c) for c in (64, 320, 640, 960, 1280, 1920, 2560, 4096)) // 8
            ws = torch.empty(need, device=self.device, dtype=torch.float64)
            self._gn_ws[nb] = ws
        return ws

    def temb_rowbias(self, timesteps: torch.Tensor) -> torch.Tensor:
        """[NB] integer timesteps -> fp32 [NB, emb_width]: all emb_layers outputs (without conv bias).
        timestep_embedding (util.py:206-230) -> time_embed MLP (openaimodel.py:340-344) -> SiLU -> emb_layers."""
        args = timesteps.to(self.device).float()[:, None] * self._freqs[None]
        t_emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1).half().contiguous()
        h = ops.linear(t_emb, self.w_te0, self.b_te0, act=ops.UDT_ACT_SILU)
        e = ops.linear(h, self.w_te2, self.b_te2, act=ops.UDT_ACT_SILU)  # SiLU of emb_layers[0] fused here
        return ops.linear(e, self.w_emb, self.b_emb, out_fp32=True)

    def context_kv(self, t_context: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[NB, L, t_context_dim] fp32 -> fp16 [NB*L, kv_width]: to_k/to_v of every t_attn layer (attention.py:144-145)."""
        nb, l, d = t_context.shape
        ctx = t_context.to(self.device).reshape(nb * l, d).half().contiguous()
        return ops.linear(ctx, self.w_kv, out=out)

    def fold_context(self, kv: torch.Tensor, ctx_len: int, first: int, count: int,
                     bufs: Optional[Dict[int, tuple]] = None) -> Dict[int, tuple]:
        """fold the context K / V of samples [first, first + count) (rows of `kv` from context_kv) into every t_attn layer's
        to_q / to_out (udt_xattn_fold; once per request — the context does not change from step to step)"""
        out: Dict[int, tuple] = {} if bufs is None else bufs
        for li, s in enumerate(self.st_layers):
            c = s.c
            npad = _round_up(s.heads * ctx_len, 64)
            if li not in out:
                out[li] = (torch.empty((count, npad, c), device=self.device, dtype=torch.float16),
                           torch.empty((count, c, npad), device=self.device, dtype=torch.float16))
            w1, w2 = out[li]
            kc = kv[first * ctx_len:, s.kv_off: s.kv_off + c]
            vc = kv[first * ctx_len:, s.kv_off + c: s.kv_off + 2 * c]
            ops.xattn_fold(kc, vc, s.w_tq, s.w_to, count, ctx_len, s.heads, 0.125, w1, w2)
        return out

    def _res(self, r: _Res, x0: torch.Tensor, x1: Optional[torch.Tensor], rowbias: torch.Tensor) -> torch.Tensor:
        nb = x0.shape[0]
        a = ops.groupnorm(x0, r.g1, r.b1, 1e-5, True, x1=x1, ws=self._ws(nb))
        rb = rowbias[:, r.emb_off: r.emb_off + r.cout]
        h = ops.conv3x3(a, r.w1, r.cb1, rowbias=rb)
        a2 = ops.groupnorm(h, r.g2, r.b2, 1e-5, True, ws=self._ws(nb))
        if r.has_skip:
            skips = [x0] if x1 is None else [x0, x1]
            return ops.conv3x3(a2, r.w2, r.cb2, skip_srcs=skips)
        return ops.conv3x3(a2, r.w2, r.cb2, residual=x0)

    def _st(self, s: _ST, x: torch.Tensor, kv: torch.Tensor, ctx_len: int, li: int) -> torch.Tensor:
        nb, hh, ww, c = x.shape
        n = hh * ww
        xn = ops.groupnorm(x, s.gn_g, s.gn_b, 1e-6, False, ws=self._ws(nb))
        t = ops.linear(xn.view(nb * n, c), s.w_in, s.b_in)
        # self-attention
        qkv = ops.linear(ops.layernorm(t, s.ln1_g, s.ln1_b), s.w_qkv)
        a = ops.fmha(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], nb, n, n, s.heads, 0.125)
        if self.skip_uc_xattn and not self.export_attn_maps and nb % 2 == 0:
            t = ops.linear(a, s.w_o1, s.b_o1, residual=t, rowbias=s.uc_rowbias(nb))   # uc half: + t_attn.to_out.bias
            hb = nb // 2
            tc = t[hb * n:]                                                            # conditional half, in place
            fold = self.xattn_fold.get(li) if self.xattn_fold is not None else None
            if fold is not None and n % 256 == 0:
                # folded form: scores = LN(t) W1[s]^T, P = softmax over the 12 tokens of every head, out = P W2[s]^T
                # (W1 / W2 = the step-invariant context K / V folded into to_q / to_out once per request)
                w1, w2 = fold
                npad = w1.shape[1]
                sc = ops.linear(ops.layernorm(tc, s.lnt_g, s.lnt_b), w1.view(hb * npad, c), groups=hb, weight_img_rows=npad)
                ops.softmax_groups(sc, s.heads, ctx_len)
                ops.linear(sc, w2.view(hb * c, npad), s.b_to, residual=tc, out=tc, groups=hb, weight_img_rows=c)
            else:
                q = ops.linear(ops.layernorm(tc, s.lnt_g, s.lnt_b), s.w_tq)
                kc = kv[hb * ctx_len:, s.kv_off: s.kv_off + c]
                vc = kv[hb * ctx_len:, s.kv_off + c: s.kv_off + 2 * c]
                a = ops.xattn_small_l(q, kc, vc, hb, n, ctx_len, s.heads, 0.125)
                ops.linear(a, s.w_to, s.b_to, residual=tc, out=tc)
            g = ops.linear(ops.layernorm(t, s.ln3_g, s.ln3_b), s.w_ff1, s.b_ff1, act=ops.UDT_ACT_GEGLU)
            t = ops.linear(g, s.w_ff2, s.b_ff2, residual=t)
            out = ops.linear(t, s.w_out, s.b_out, residual=x.view(nb * n, c))
            return out.view(nb, hh, ww, c)
        t = ops.linear(a, s.w_o1, s.b_o1, residual=t)
        # textual cross-attention
        q = ops.linear(ops.layernorm(t, s.lnt_g, s.lnt_b), s.w_tq)
        probs = None
        if self.export_attn_maps:
            probs = torch.empty((nb * s.heads, n, ctx_len), device=self.device, dtype=torch.float32)
            self.attn_map_cache[li]["size"] = int(n ** 0.5)
            self.attn_map_cache[li]["attn_map"] = probs
        kc = kv[:, s.kv_off: s.kv_off + c]
        vc = kv[:, s.kv_off + c: s.kv_off + 2 * c]
        a = ops.xattn_small_l(q, kc, vc, nb, n, ctx_len, s.heads, 0.125, probs=probs)
        t = ops.linear(a, s.w_to, s.b_to, residual=t)
        # GEGLU feed-forward
        g = ops.linear(ops.layernorm(t, s.ln3_g, s.ln3_b), s.w_ff1, s.b_ff1, act=ops.UDT_ACT_GEGLU)
        t = ops.linear(g, s.w_ff2, s.b_ff2, residual=t)
        out = ops.linear(t, s.w_out, s.b_out, residual=x.view(nb * n, c))
        return out.view(nb, hh, ww, c)

    def _run(self, layers: list, h: torch.Tensor, skip: Optional[torch.Tensor], rowbias, kv, ctx_len, st_counter) -> torch.Tensor:
        for layer in layers:
            kind = layer[0]
            if kind == "res":
                h = self._res(layer[1], h, skip, rowbias)
                skip = None
            elif kind == "st":
                h = self._st(layer[1], h, kv, ctx_len, st_counter[0])
                st_counter[0] += 1
            elif kind == "down":  # conv3x3 stride 2 pad 1 (openaimodel.py:132-139): strided TMA boxes
                _, w, b, c = layer
                h = ops.conv3x3(h, w, b, stride=2, pad=1)
            elif kind == "up":  # nearest x2 + conv3x3 (openaimodel.py:99-102)
                _, w, b, c, w4 = layer
                if h.shape[0] * h.shape[1] * h.shape[2] >= self.up2_min_rows:
                    h = ops.conv3x3_up2(h, w4, b)        # four 2x2 phase convs on the low-res tensor
                else:                                     # tiny M: four launches cost more than the FLOPs they save
                    h = ops.conv3x3(ops.upsample2x(h), w, b)
            elif kind == "conv_in":  # Cin = 9 stored as 16 channels; TMA zero-fills each tap up to 64
                h = ops.conv3x3(h, self.w_conv_in, self.b_conv_in)
        return h

    # ------------------------------------------------------------------------------------------ forward
    def forward_nhwc(self, x: torch.Tensor, rowbias: torch.Tensor, kv: torch.Tensor, ctx_len: int,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: fp16 [NB, H, W, cin_pad]; rowbias: temb_rowbias(t); kv: context_kv(ctx) -> fp32 [NB, H, W, out_channels]"""
        if not self.export_attn_maps:
            for item in self.attn_map_cache:
                item["attn_map"] = None
        hs: List[torch.Tensor] = []
        cnt = [0]
        h = x
        for layers in self.input_plan:
            h = self._run(layers, h, None, rowbias, kv, ctx_len, cnt)
            hs.append(h)
        h = self._run(self.middle_plan, h, None, rowbias, kv, ctx_len, cnt)
        for layers in self.output_plan:
            h = self._run(layers, h, hs.pop(), rowbias, kv, ctx_len, cnt)
        a = ops.groupnorm(h, self.out_g, self.out_b, 1e-5, True, ws=self._ws(h.shape[0]))
        return ops.conv3x3(a, self.w_out, self.b_out, out=out, out_fp32=True)

    def forward(self, x: torch.Tensor, timesteps: torch.Tensor, t_context: torch.Tensor) -> torch.Tensor:
        """Reference-shaped call (openaimodel.py:593): x fp32 NCHW [NB, in_channels, H, W] -> fp32 NCHW."""
        xh = ops.nchw_to_nhwc_f16(x.to(self.device).float().contiguous(), cpad=self.cin_pad)
        y = self.forward_nhwc(xh, self.temb_rowbias(timesteps), self.context_kv(t_context), t_context.shape[1])
        return ops.nhwc_to_nchw_f32(y, self.out_channels)
