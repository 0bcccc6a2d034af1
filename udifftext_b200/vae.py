"""AutoencoderKL encoder / decoder executed on the sm_100a kernels (reference: sgm/modules/diffusionmodules/model.py
:91-148 ResnetBlock, :201-262 MemoryEfficientAttnBlock, :482-596 Encoder, :599-743 Decoder; sgm/models/autoencoder.py
:282-316 quant / post_quant convs).

Same conventions as `unet.py`: NHWC fp16 activations, fp32 accumulation / statistics, every tensor op is a C-ABI
launch.  Fusions relative to the reference graph:
  * GroupNorm(eps 1e-6)+swish is one op; conv bias, the 1x1 `nin_shortcut` (extra K segment) and the residual add
    are the epilogue of the 3x3 implicit GEMM;
  * the single-head mid attention (d = C = 512) runs as GEMM (q k^T, 1/sqrt(C) folded into the q projection) ->
    row softmax -> GEMM (P V); V is produced already transposed ([C, N], the K-major operand the second GEMM
    needs) by swapping the operand roles of its projection GEMM, its bias moves behind P V (softmax rows sum to 1);
  * `quant_conv` (1x1, 8->8) is folded into the encoder's conv_out weights; `post_quant_conv` + the 1/scale_factor
    of `decode_first_stage` is one per-pixel affine kernel; the final `clamp((x+1)/2, 0, 1)` (test.py:38) is fused
    into the NHWC->NCHW conversion of the decoder output.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops, pack

SD = Dict[str, torch.Tensor]


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _VRes:
    """ResnetBlock weights (model.py:91-148, temb_channels = 0)"""

    def __init__(self, sd: SD, p: str, dev):
        f = lambda k: pack.f32(sd[p + k]).to(dev)
        self.g1, self.b1 = f("norm1.weight"), f("norm1.bias")
        self.g2, self.b2 = f("norm2.weight"), f("norm2.bias")
        self.w1 = pack.pack_conv3x3(sd[p + "conv1.weight"]).to(dev)
        self.cb1 = f("conv1.bias")
        cb2 = sd[p + "conv2.bias"].float()
        self.has_skip = (p + "nin_shortcut.weight") in sd
        if self.has_skip:
            self.w2 = pack.pack_conv3x3(sd[p + "conv2.weight"], [sd[p + "nin_shortcut.weight"]]).to(dev)
            cb2 = cb2 + sd[p + "nin_shortcut.bias"].float()
        else:
            self.w2 = pack.pack_conv3x3(sd[p + "conv2.weight"]).to(dev)
        self.cb2 = cb2.contiguous().to(dev)


class _VAttn:
    """single-head attention over pixels (model.py:201-262)"""

    def __init__(self, sd: SD, p: str, dev):
        c = sd[p + "q.weight"].shape[0]
        self.c = c
        s = float(c) ** -0.5
        self.g, self.b = pack.f32(sd[p + "norm.weight"]).to(dev), pack.f32(sd[p + "norm.bias"]).to(dev)
        wq = sd[p + "q.weight"].float().reshape(c, c) * s
        wk = sd[p + "k.weight"].float().reshape(c, c)
        self.w_qk = torch.cat([wq, wk], dim=0).half().contiguous().to(dev)
        self.b_qk = torch.cat([sd[p + "q.bias"].float() * s, sd[p + "k.bias"].float()]).contiguous().to(dev)
        self.w_v = pack.pack_linear(sd[p + "v.weight"]).to(dev)          # used as the A operand: V^T = Wv hn^T
        self.b_v = pack.f32(sd[p + "v.bias"]).to(dev)                     # added after P V
        self.w_o = pack.pack_linear(sd[p + "proj_out.weight"]).to(dev)
        self.b_o = pack.f32(sd[p + "proj_out.bias"]).to(dev)
        self._wv_tiled: dict = {}

    def w_v_tiled(self, nb: int) -> torch.Tensor:
        """W_v stacked nb times: the A operand of the batched V^T GEMM (rows [i*c, (i+1)*c) multiply image i's pixels)"""
        t = self._wv_tiled.get(nb)
        if t is None:
            t = self.w_v.repeat(nb, 1).contiguous()
            self._wv_tiled[nb] = t
        return t


class VAEB200:
    """Inference-only AutoencoderKL on B200.  `sd` keys are relative to the autoencoder root
    (`first_stage_model.` or `conditioner.embedders.2.model.`)."""

    def __init__(self, sd: SD, device, ch: int = 128, ch_mult: Sequence[int] = (1, 2, 4, 4), num_res_blocks: int = 2,
                 z_channels: int = 4, in_channels: int = 3, out_ch: int = 3, build_encoder: bool = True,
                 build_decoder: bool = True, **_ignored):
        dev = torch.device(device)
        self.device = dev
        self.z_channels, self.in_channels, self.out_ch = z_channels, in_channels, out_ch
        self.levels = len(ch_mult)
        self._gn_ws: Optional[torch.Tensor] = None
        self._gn_ws_key = (0, 0)
        self._graphs: dict = {}
        self._gn_ws_old: list = []
        if build_encoder and "encoder.conv_in.weight" in sd:
            self._build_encoder(sd, dev, num_res_blocks)
        if build_decoder and "decoder.conv_in.weight" in sd:
            self._build_decoder(sd, dev, num_res_blocks)

    # ------------------------------------------------------------------------------------------ weights
    def _build_encoder(self, sd: SD, dev, nrb: int) -> None:
        w0 = sd["encoder.conv_in.weight"].float()
        self.e_cin_pad = _round_up(w0.shape[1], 8)
        self.e_w_in = pack.pack_conv3x3(w0, cin_pad=self.e_cin_pad).to(dev)
        self.e_b_in = pack.f32(sd["encoder.conv_in.bias"]).to(dev)
        self.e_down: List[dict] = []
        for lvl in range(self.levels):
            blocks = [_VRes(sd, f"encoder.down.{lvl}.block.{i}.", dev) for i in range(nrb)]
            item = {"blocks": blocks, "down": None}
            k = f"encoder.down.{lvl}.downsample.conv."
            if k + "weight" in sd:
                item["down"] = (pack.pack_conv3x3(sd[k + "weight"]).to(dev), pack.f32(sd[k + "bias"]).to(dev))
            self.e_down.append(item)
        self.e_mid1 = _VRes(sd, "encoder.mid.block_1.", dev)
        self.e_attn = _VAttn(sd, "encoder.mid.attn_1.", dev)
        self.e_mid2 = _VRes(sd, "encoder.mid.block_2.", dev)
        self.e_g, self.e_b = pack.f32(sd["encoder.norm_out.weight"]).to(dev), pack.f32(sd["encoder.norm_out.bias"]).to(dev)
        # quant_conv (1x1) folded into conv_out: W' = Wq Wc, b' = Wq bc + bq (exact in real arithmetic; combined in fp32)
        wc = sd["encoder.conv_out.weight"].float()
        bc = sd["encoder.conv_out.bias"].float()
        wq = sd["quant_conv.weight"].float().reshape(sd["quant_conv.weight"].shape[0], -1)
        wcomb = torch.einsum("om,mikl->oikl", wq, wc)
        self.e_w_out = pack.pack_conv3x3(wcomb).to(dev)
        self.e_b_out = (wq @ bc + sd["quant_conv.bias"].float()).contiguous().to(dev)
        self.moment_channels = wq.shape[0]

    def _build_decoder(self, sd: SD, dev, nrb: int) -> None:
        zc = self.z_channels
        self.d_wpq = pack.f32(sd["post_quant_conv.weight"].reshape(zc, zc)).to(dev)
        self.d_bpq = pack.f32(sd["post_quant_conv.bias"]).to(dev)
        self.d_cin_pad = _round_up(zc, 8)
        self.d_w_in = pack.pack_conv3x3(sd["decoder.conv_in.weight"].float(), cin_pad=self.d_cin_pad).to(dev)
        self.d_b_in = pack.f32(sd["decoder.conv_in.bias"]).to(dev)
        self.d_mid1 = _VRes(sd, "decoder.mid.block_1.", dev)
        self.d_attn = _VAttn(sd, "decoder.mid.attn_1.", dev)
        self.d_mid2 = _VRes(sd, "decoder.mid.block_2.", dev)
        self.d_up: List[dict] = []
        for lvl in range(self.levels):
            blocks = [_VRes(sd, f"decoder.up.{lvl}.block.{i}.", dev) for i in range(nrb + 1)]
            item = {"blocks": blocks, "up": None}
            k = f"decoder.up.{lvl}.upsample.conv."
            if k + "weight" in sd:
                item["up"] = ([t.to(dev) for t in pack.pack_conv3x3_up2(sd[k + "weight"])], pack.f32(sd[k + "bias"]).to(dev))
            self.d_up.append(item)
        self.d_g, self.d_b = pack.f32(sd["decoder.norm_out.weight"]).to(dev), pack.f32(sd["decoder.norm_out.bias"]).to(dev)
        self.d_w_out = pack.pack_conv3x3(sd["decoder.conv_out.weight"]).to(dev)
        self.d_b_out = pack.f32(sd["decoder.conv_out.bias"]).to(dev)

    # ------------------------------------------------------------------------------------------ pieces
    def _ws(self, nb: int, hw: int) -> torch.Tensor:
        if self._gn_ws is None or self._gn_ws_key[0] < nb or self._gn_ws_key[1] < hw:
            nb_, hw_ = max(nb, self._gn_ws_key[0]), max(hw, self._gn_ws_key[1])
            need = max(ops.groupnorm_ws_bytes(nb_, hw_, c) for c in (64, 128, 256, 512, 1024)) // 8
            if self._gn_ws is not None:
                self._gn_ws_old.append(self._gn_ws)        # captured graphs may still point at the smaller workspace
            self._gn_ws = torch.empty(need, device=self.device, dtype=torch.float64)
            self._gn_ws_key = (nb_, hw_)
        return self._gn_ws

    def _gn(self, x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, silu: bool) -> torch.Tensor:
        nb, hh, ww, _ = x.shape
        return ops.groupnorm(x, g, b, 1e-6, silu, ws=self._ws(nb, hh * ww))

    def _res(self, r: _VRes, x: torch.Tensor) -> torch.Tensor:
        h = ops.conv3x3(self._gn(x, r.g1, r.b1, True), r.w1, r.cb1)
        a2 = self._gn(h, r.g2, r.b2, True)
        if r.has_skip:
            return ops.conv3x3(a2, r.w2, r.cb2, skip_srcs=[x])
        return ops.conv3x3(a2, r.w2, r.cb2, residual=x)

    def _attn(self, a: _VAttn, x: torch.Tensor) -> torch.Tensor:
        """single-head attention over the pixels (model.py:228-262), all images of the batch per launch: the per-image
        operands K_i and V_i^T enter udt_igemm as per-image WEIGHTS (weight_img_rows), so q k^T, the row softmax, V^T, P V and
        proj_out are five launches for the whole batch (the N x N probabilities still pass through HBM as fp16 — d = 512
        does not fit the TMEM budget of the fused d = 64 kernel)"""
        nb, hh, ww, c = x.shape
        n = hh * ww
        hn = self._gn(x, a.g, a.b, False).view(nb * n, c)
        qk = ops.linear(hn, a.w_qk, a.b_qk)                       # [nb*n, 2c]: q (pre-scaled) | k
        xf = x.view(nb * n, c)
        if n % 256 or c % 256 or nb == 1:                         # per-image weights need whole CTA-pair tiles per image (256 rows)
            out = torch.empty((nb * n, c), device=x.device, dtype=torch.float16)
            for i in range(nb):
                rows = slice(i * n, (i + 1) * n)
                s = ops.linear(qk[rows, :c], qk[rows, c:])        # q k^T  [n, n]
                ops.softmax_rows_(s, 1.0)
                vt = ops.linear(a.w_v, hn[rows])                  # V^T (without bias) [c, n]
                o = ops.linear(s, vt, a.b_v)                      # P V + b_v  [n, c]
                ops.linear(o, a.w_o, a.b_o, residual=xf[rows], out=out[rows])
            return out.view(nb, hh, ww, c)
        s = ops.linear(qk[:, :c], qk[:, c:], groups=nb, weight_img_rows=n)          # q_i k_i^T  [nb*n, n]
        ops.softmax_rows_(s, 1.0)
        wv = a.w_v_tiled(nb)                                                        # W_v repeated per image  [nb*c, c]
        vt = ops.linear(wv, hn, groups=nb, weight_img_rows=n)                       # V_i^T (without bias)  [nb*c, n]
        o = ops.linear(s, vt, a.b_v, groups=nb, weight_img_rows=c)                  # P_i V_i + b_v  [nb*n, c]
        return ops.linear(o, a.w_o, a.b_o, residual=xf).view(nb, hh, ww, c)

    # ------------------------------------------------------------------------------------------ encode
    def encode_moments_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 NCHW [B, 3, H, W] -> moments fp32 NHWC [B, H/8, W/8, 2*z] (Encoder.forward + quant_conv)"""
        nb, _, hh, ww = x.shape
        xh = ops.nchw_to_nhwc_f16(x.to(self.device).float().contiguous(), cpad=self.e_cin_pad)
        h = ops.conv3x3(xh, self.e_w_in, self.e_b_in)
        for item in self.e_down:
            for r in item["blocks"]:
                h = self._res(r, h)
            if item["down"] is not None:                           # pad (0,1,0,1) + conv3x3 stride 2 (model.py:77-85)
                w, b = item["down"]
                h = ops.conv3x3(h, w, b, stride=2, pad=0)
        h = self._res(self.e_mid1, h)
        h = self._attn(self.e_attn, h)
        h = self._res(self.e_mid2, h)
        a = self._gn(h, self.e_g, self.e_b, True)
        return ops.conv3x3(a, self.e_w_out, self.e_b_out, out_fp32=True)

    def graphed(self, name: str):
        """CUDA-graphed twin of `encode_moments_nhwc` / `decode` (udifftext_b200.graphs.GraphCache): same kernels, one graph
        launch per call; returns static buffers (valid until the next call of the same shape)"""
        from .graphs import GraphCache
        g = self._graphs.get(name)
        if g is None:
            g = GraphCache(getattr(self, name))
            self._graphs[name] = g
        return g

    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """reference-shaped: fp32 NCHW moments [B, 2*z, h, w] (autoencoder.py:304-309 before the posterior)"""
        m = self.encode_moments_nhwc(x)
        return ops.nhwc_to_nchw_f32(m, self.moment_channels)

    # ------------------------------------------------------------------------------------------ decode
    def decode_nhwc(self, z: torch.Tensor, in_scale: float = 1.0) -> torch.Tensor:
        """z fp32 NCHW [B, z, h, w] -> fp32 NHWC [B, 8h, 8w, 4] (3 valid channels)"""
        nb, _, hh, ww = z.shape
        zp = ops.pointwise_affine(z.to(self.device).float().contiguous(), self.d_wpq, self.d_bpq, self.d_cin_pad, in_scale)
        h = ops.conv3x3(zp, self.d_w_in, self.d_b_in)
        h = self._res(self.d_mid1, h)
        h = self._attn(self.d_attn, h)
        h = self._res(self.d_mid2, h)
        for lvl in reversed(range(self.levels)):
            item = self.d_up[lvl]
            for r in item["blocks"]:
                h = self._res(r, h)
            if item["up"] is not None:                             # nearest x2 + conv3x3 (model.py:55-68)
                w4, b = item["up"]
                h = ops.conv3x3_up2(h, w4, b)                                 # four 2x2 phase convs on the low-res tensor
        a = self._gn(h, self.d_g, self.d_b, True)
        nb, oh, ow, _ = a.shape
        out = torch.empty((nb, oh, ow, 4), device=self.device, dtype=torch.float32)
        ops.igemm([(a, a.shape[-1], a.shape[-1], 9)], nb, oh, ow, self.d_w_out, self.out_ch, out, 4, bias=self.d_b_out,
                  out_fp32=True)
        return out

    def decode(self, z: torch.Tensor, in_scale: float = 1.0, out_scale: float = 1.0, out_shift: float = 0.0,
               clamp01: bool = False, chunk: int = 8) -> torch.Tensor:
        """reference-shaped: fp32 NCHW [B, 3, 8h, 8w] = out_scale * Decoder(post_quant_conv(z * in_scale)) + out_shift
        (optionally clamped to [0, 1]); decoded in chunks of `chunk` images to bound activation memory."""
        nb, _, hh, ww = z.shape
        out = torch.empty((nb, self.out_ch, 8 * hh, 8 * ww), device=self.device, dtype=torch.float32)
        for i in range(0, nb, chunk):
            y = self.decode_nhwc(z[i: i + chunk], in_scale)
            ops.nhwc_to_nchw_f32(y, self.out_ch, out_scale, out_shift, clamp01, out=out[i: i + chunk])
        return out
