"""Thin torch-tensor front end over the C-ABI (udifftext_b200.lib).

torch is used only for device memory and the current stream; every op below is one or two launches of a
hand-written sm_100a kernel.  Activations are fp16, channels-last (`[NB, H, W, C]` or `[rows, C]`).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import lib as _lib
from .lib import UDT_ACT_GEGLU, UDT_ACT_NONE, UDT_ACT_SILU, GemmSrc  # noqa: F401


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous CUDA {dtype} tensor, got {t.dtype} {t.device} "
                         f"contiguous={t.is_contiguous()}")


def igemm(
    srcs: Sequence[Tuple[torch.Tensor, int, int, int]],
    nb: int, h: int, w: int,
    weight: torch.Tensor,
    n_out: int,
    out: torch.Tensor,
    ldo: int,
    bias: Optional[torch.Tensor] = None,
    rowbias: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    ldr: int = 0,
    ld_rowbias: int = 0,
    out_fp32: bool = False,
    act: int = UDT_ACT_NONE,
    bn_hint: int = 0,
) -> torch.Tensor:
    """Segmented implicit GEMM (udt_igemm).  `srcs` = [(tensor, C, ld, taps), ...]."""
    L = _lib.load()
    arr = (GemmSrc * len(srcs))()
    for i, (t, c, ld, taps) in enumerate(srcs):
        arr[i].ptr = t.data_ptr()
        arr[i].C = c
        arr[i].ld = ld
        arr[i].taps = taps
    rc = L.udt_igemm(arr, len(srcs), nb, h, w, weight.data_ptr(), n_out, _ptr(bias), _ptr(rowbias),
                     (ld_rowbias or n_out) if rowbias is not None else 0, _ptr(residual), ldr,
                     out.data_ptr(), ldo, int(out_fp32), act, bn_hint, _stream())
    _lib.check(rc, "udt_igemm")
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, act: int = UDT_ACT_NONE,
           out_fp32: bool = False, bn_hint: int = 0) -> torch.Tensor:
    """y[M, N] = act(x[M, K] @ weight[N, K]^T + bias) (+ residual); fp16 in, fp16 (or fp32) out."""
    m, k = x.shape
    n = weight.shape[0]
    n_log = n // 2 if act == UDT_ACT_GEGLU else n
    if out is None:
        out = torch.empty((m, n_log), device=x.device, dtype=torch.float32 if out_fp32 else torch.float16)
    return igemm([(x, k, x.stride(0), 1)], 1, 1, m, weight, n, out, out.stride(0), bias=bias, residual=residual,
                 ldr=0 if residual is None else residual.stride(0), out_fp32=out_fp32, act=act, bn_hint=bn_hint)


def conv3x3(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
            rowbias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
            skip_srcs: Sequence[torch.Tensor] = (), out: Optional[torch.Tensor] = None, out_fp32: bool = False,
            bn_hint: int = 0) -> torch.Tensor:
    """3x3 / stride 1 / pad 1 conv on NHWC fp16 with packed weight [Cout, 9*Cin (+ sum skip C)];
    `skip_srcs` are extra NHWC tensors consumed point-wise (fused 1x1 skip connection)."""
    nb, h, w, c = x.shape
    n = weight.shape[0]
    if out is None:
        out = torch.empty((nb, h, w, n), device=x.device, dtype=torch.float32 if out_fp32 else torch.float16)
    srcs = [(x, c, c, 9)] + [(s, s.shape[-1], s.shape[-1], 1) for s in skip_srcs]
    return igemm(srcs, nb, h, w, weight, n, out, n, bias=bias, rowbias=rowbias, residual=residual,
                 ldr=0 if residual is None else residual.shape[-1], out_fp32=out_fp32, bn_hint=bn_hint,
                 ld_rowbias=0 if rowbias is None else rowbias.stride(0))


def groupnorm_ws_bytes(nb: int, hw: int, c: int, groups: int = 32) -> int:
    return int(_lib.load().udt_groupnorm_ws_bytes(nb, hw, c, groups))


def groupnorm(x0: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, silu: bool,
              x1: Optional[torch.Tensor] = None, groups: int = 32, out: Optional[torch.Tensor] = None,
              ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm(+SiLU) over NHWC fp16; with `x1` normalises cat([x0, x1], channel) and writes the concat."""
    L = _lib.load()
    nb = x0.shape[0]
    c0 = x0.shape[-1]
    c1 = 0 if x1 is None else x1.shape[-1]
    hw = x0.numel() // (nb * c0)
    if out is None:
        out = torch.empty(tuple(x0.shape[:-1]) + (c0 + c1,), device=x0.device, dtype=torch.float16)
    if ws is None or ws.numel() * 8 < groupnorm_ws_bytes(nb, hw, c0 + c1, groups):
        ws = torch.empty(groupnorm_ws_bytes(nb, hw, c0 + c1, groups) // 8, device=x0.device, dtype=torch.float64)
    rc = L.udt_groupnorm_nhwc(x0.data_ptr(), c0, _ptr(x1), c1, out.data_ptr(), nb, hw, groups, gamma.data_ptr(),
                              beta.data_ptr(), float(eps), int(silu), ws.data_ptr(), _stream())
    _lib.check(rc, "udt_groupnorm_nhwc")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    L = _lib.load()
    c = x.shape[-1]
    rows = x.numel() // c
    if out is None:
        out = torch.empty_like(x)
    rc = L.udt_layernorm(x.data_ptr(), out.data_ptr(), rows, c, gamma.data_ptr(), beta.data_ptr(), float(eps), _stream())
    _lib.check(rc, "udt_layernorm")
    return out


def fmha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, b: int, nq: int, nkv: int, heads: int, scale: float,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q/k/v: 2-D fp16 views [B*N, >= heads*64] (may be column slices of one fused QKV buffer)."""
    L = _lib.load()
    if out is None:
        out = torch.empty((b * nq, heads * 64), device=q.device, dtype=torch.float16)
    rc = L.udt_fmha_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), b, nq, nkv, heads, q.stride(0),
                        k.stride(0), v.stride(0), out.stride(0), float(scale), _stream())
    _lib.check(rc, "udt_fmha_fwd")
    return out


def xattn_small_l(q: torch.Tensor, kc: torch.Tensor, vc: torch.Tensor, b: int, n: int, l: int, heads: int,
                  scale: float, out: Optional[torch.Tensor] = None, probs: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B*N, heads*64]; kc/vc [B*L, heads*64] (row pitch may be larger); probs fp32 [B*heads, N, L] optional."""
    L = _lib.load()
    if out is None:
        out = torch.empty((b * n, heads * 64), device=q.device, dtype=torch.float16)
    rc = L.udt_xattn_small_l(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), out.data_ptr(), _ptr(probs), b, n, l, heads,
                             q.stride(0), kc.stride(0), out.stride(0), float(scale), _stream())
    _lib.check(rc, "udt_xattn_small_l")
    return out


def softmax_rows_(x: torch.Tensor, scale: float) -> torch.Tensor:
    L = _lib.load()
    rows, cols = x.shape
    _lib.check(L.udt_softmax_rows(x.data_ptr(), rows, cols, x.stride(0), float(scale), _stream()), "udt_softmax_rows")
    return x


def cfg_pack(x: torch.Tensor, cat_uc: torch.Tensor, cat_c: torch.Tensor, c_in: float, out: torch.Tensor) -> torch.Tensor:
    L = _lib.load()
    b = x.shape[0]
    hw = x.numel() // (b * 4)
    _lib.check(L.udt_cfg_pack(x.data_ptr(), cat_uc.data_ptr(), cat_c.data_ptr(), out.data_ptr(), b, hw, float(c_in),
                              _stream()), "udt_cfg_pack")
    return out


def cfg_euler_step_(x: torch.Tensor, eps2b: torch.Tensor, cfg_scale: float, dsigma: float) -> torch.Tensor:
    L = _lib.load()
    b = x.shape[0]
    hw = x.numel() // (b * 4)
    _lib.check(L.udt_cfg_euler_step(x.data_ptr(), eps2b.data_ptr(), b, hw, float(cfg_scale), float(dsigma), _stream()),
               "udt_cfg_euler_step")
    return x


def upsample2x(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    L = _lib.load()
    nb, h, w, c = x.shape
    if out is None:
        out = torch.empty((nb, 2 * h, 2 * w, c), device=x.device, dtype=torch.float16)
    _lib.check(L.udt_upsample2x_nhwc(x.data_ptr(), out.data_ptr(), nb, h, w, c, _stream()), "udt_upsample2x_nhwc")
    return out


def im2col3x3(x: torch.Tensor, stride: int, pad_lo: int, ho: int, wo: int, kpad: int,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    L = _lib.load()
    nb, h, w, c = x.shape
    if out is None:
        out = torch.empty((nb * ho * wo, kpad), device=x.device, dtype=torch.float16)
    _lib.check(L.udt_im2col3x3_nhwc(x.data_ptr(), out.data_ptr(), nb, h, w, c, x.stride(2), stride, pad_lo, ho, wo, kpad,
                                    _stream()), "udt_im2col3x3_nhwc")
    return out


def nchw_to_nhwc_f16(x: torch.Tensor, cpad: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    L = _lib.load()
    nb, c, h, w = x.shape
    cpad = c if cpad is None else cpad
    if out is None:
        out = torch.empty((nb, h, w, cpad), device=x.device, dtype=torch.float16)
    _lib.check(L.udt_nchw_f32_to_nhwc_f16(x.data_ptr(), out.data_ptr(), nb, c, h * w, cpad, _stream()),
               "udt_nchw_f32_to_nhwc_f16")
    return out


def nhwc_to_nchw_f32(x: torch.Tensor, c: int, scale: float = 1.0, shift: float = 0.0, clamp01: bool = False,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    L = _lib.load()
    nb, h, w, ld = x.shape
    if out is None:
        out = torch.empty((nb, c, h, w), device=x.device, dtype=torch.float32)
    _lib.check(L.udt_nhwc_to_nchw_f32(x.data_ptr(), int(x.dtype == torch.float32), out.data_ptr(), nb, c, h * w, ld,
                                      float(scale), float(shift), int(clamp01), _stream()), "udt_nhwc_to_nchw_f32")
    return out
