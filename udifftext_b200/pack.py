"""Weight repacking: reference `state_dict` layouts (conv OIHW fp32, Linear [out, in] fp32) -> the K-major fp16
layouts udt_igemm consumes.  Pure tensor reshuffles, run once after checkpoint load (any device)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

GEGLU_TILE = 256  # must equal udt_geglu_tile(); passed to udt_igemm as bn_hint (128 is also accepted)


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def _pad_k(w2d: torch.Tensor) -> torch.Tensor:
    """[O, K] -> [O, round_up(K, 64)] zero padded (TMA zero-fills the activation side of the same K block)"""
    o, k = w2d.shape
    if k % 64 == 0:
        return w2d
    out = w2d.new_zeros((o, _pad64(k)))
    out[:, :k] = w2d
    return out


def pack_conv3x3(w_oihw: torch.Tensor, skip_1x1: Sequence[torch.Tensor] = (), cin_pad: Optional[int] = None) -> torch.Tensor:
    """[O, I, 3, 3] -> fp16 [O, 9*I64 (+ sum of skip I64)] with K ordered (tap = ky*3 + kx, channel) and every
    tap's channel block zero padded to a multiple of 64 (I64); optional 1x1 skip-connection weights
    ([O, I_s, 1, 1] or [O, I_s]) are appended as extra K segments.  `cin_pad` (>= I) is the channel count of the
    activation tensor when it is stored wider than I (zero channels)."""
    o, i, kh, kw = w_oihw.shape
    assert (kh, kw) == (3, 3)
    i64 = _pad64(cin_pad if cin_pad is not None else i)
    w = w_oihw.new_zeros((o, 3, 3, i64))
    w[..., :i] = w_oihw.permute(0, 2, 3, 1)
    parts = [w.reshape(o, 9 * i64)]
    for s in skip_1x1:
        parts.append(_pad_k(s.reshape(o, -1)))
    return torch.cat(parts, dim=1).to(torch.float16).contiguous()


def pack_conv3x3_up2(w_oihw: torch.Tensor) -> list:
    """nearest-2x upsample followed by a 3x3 conv (pad 1) == four 2x2 convs on the LOW-resolution input, one per output
    phase (py, px): output pixel (2h+py, 2w+px) reads source rows {h-1, h} (py = 0) or {h, h+1} (py = 1), and the 3x3 taps
    that land on the same source pixel are summed (in fp32, then rounded to fp16).  Returns the 4 packed weights
    [O, 4*I64] in phase order (0,0), (0,1), (1,0), (1,1), K ordered (tap = ty*2 + tx, channel padded to 64)."""
    o, i, kh, kw = w_oihw.shape
    assert (kh, kw) == (3, 3)
    w = w_oihw.float()
    i64 = _pad64(i)
    taps = {0: ([0], [1, 2]), 1: ([0, 1], [2])}     # phase -> (3x3 indices feeding window slot 0, slot 1)
    out = []
    for py in (0, 1):
        for px in (0, 1):
            wp = w.new_zeros((o, 2, 2, i64))
            for ty in (0, 1):
                for tx in (0, 1):
                    acc = 0
                    for ky in taps[py][ty]:
                        for kx in taps[px][tx]:
                            acc = acc + w[:, :, ky, kx]
                    wp[:, ty, tx, :i] = acc
            out.append(wp.reshape(o, 4 * i64).to(torch.float16).contiguous())
    return out


def pack_linear(w: torch.Tensor) -> torch.Tensor:
    """[out, in] (or 1x1 conv [O, I, 1, 1]) -> fp16 [out, round_up(in, 64)]"""
    return _pad_k(w.reshape(w.shape[0], -1)).to(torch.float16).contiguous()


def pack_geglu(w: torch.Tensor, b: Optional[torch.Tensor]) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """GEGLU projection [2*inner, in]: rows [0, inner) are x, [inner, 2*inner) the gate (attention.py:49-51).
    Interleave per column tile of GEGLU_TILE: tile t = [x rows t*T/2 .. (t+1)*T/2 | gate rows ...]."""
    two_inner, k = w.shape
    inner = two_inner // 2
    half = GEGLU_TILE // 2
    assert inner % half == 0, f"GEGLU inner dim {inner} must be a multiple of {half}"
    wx = w[:inner].reshape(inner // half, half, k)
    wg = w[inner:].reshape(inner // half, half, k)
    wp = torch.cat([wx, wg], dim=1).reshape(two_inner, k).to(torch.float16).contiguous()
    bp = None
    if b is not None:
        bx = b[:inner].reshape(inner // half, half)
        bg = b[inner:].reshape(inner // half, half)
        bp = torch.cat([bx, bg], dim=1).reshape(two_inner).to(torch.float32).contiguous()
    return wp, bp


def f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()
