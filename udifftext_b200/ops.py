"""Thin torch-tensor front end over the C-ABI (udifftext_b200.lib).

torch is used only for device memory and the current stream; every op below is one or two launches of a
hand-written sm_100a kernel.  Activations are fp16, channels-last (`[NB, H, W, C]` or `[rows, C]`).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence, Tuple

import torch

from . import lib as _lib
from .lib import UDT_ACT_GEGLU, UDT_ACT_GELU, UDT_ACT_NONE, UDT_ACT_RELU, UDT_ACT_SILU, GemmSrc, IGemmDesc  # noqa: F401


_launches = 0   # C-ABI compute calls issued by this process (each is one or two kernel launches of libudt_b200)
SHAPE_LOG = None  # when a list: one problem-shape tuple per call (scripts/profile_unet_step.py)
_prof = None    # when a list: (entry point, start event, end event) per call — see profile_step()
CALL_LOG = None  # when a list: (entry point, bound function, args, shape) per call, for scripts/profile_step_graph.py
KEY_LOG = None   # when a list: the tuning key of every udt_igemm call (scripts/tune_igemm_bn.py)
IGEMM_TUNING = None   # {problem key: column tile}: measured exceptions to the library's cost model (udifftext_b200/tuning/)


def _igemm_tuning() -> dict:
    """lazily load the measured BN table (scripts/tune_igemm_bn.py); UDT_IGEMM_TUNING=0 ignores it"""
    global IGEMM_TUNING
    if IGEMM_TUNING is None:
        IGEMM_TUNING = {}
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuning", "igemm_bn_b200.json")
        if os.environ.get("UDT_IGEMM_TUNING", "1") != "0" and os.path.exists(path):
            import json
            with open(path) as f:
                IGEMM_TUNING = json.load(f)
    return IGEMM_TUNING


def launch_count() -> int:
    return _launches


def count_launches(n: int) -> None:
    """account for kernels launched by a CUDA-graph replay (udifftext_b200.graphs)"""
    global _launches
    _launches += int(n)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _invoke(name: str, *args, shape=None) -> None:
    """call one C-ABI entry point on torch's current stream; raises UdtError on a non-zero return code"""
    global _launches
    if SHAPE_LOG is not None:
        SHAPE_LOG.append(shape if shape is not None else tuple(a for a in args if isinstance(a, int) and a < (1 << 24)))
    fn = getattr(_lib.load(), name)
    _launches += 1
    if CALL_LOG is not None:
        CALL_LOG.append((name, fn, args, shape if shape is not None else tuple(a for a in args if isinstance(a, int) and a < (1 << 24))))
    if _prof is None:
        rc = fn(*args, _stream())
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args, _stream())
        e1.record()
        _prof.append((name, e0, e1))
    _lib.check(rc, name)


def profile_callable(fn, reps: int = 20) -> dict:
    """device time of every C-ABI call `fn()` makes: the calls are recorded during one eager execution, then each distinct
    (entry point, problem shape) is replayed `reps` times back to back inside a CUDA graph and timed with CUDA events"""
    global CALL_LOG, SHAPE_LOG
    keep = []                       # keep the logged step's temporaries alive: the recorded raw pointers stay valid
    orig_empty = torch.empty

    def empty_keep(*a, **k):
        t = orig_empty(*a, **k)
        keep.append(t)
        return t

    torch.empty = empty_keep
    shape_log_prev = SHAPE_LOG
    SHAPE_LOG, CALL_LOG = [], []
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        torch.empty = orig_empty
        calls, CALL_LOG, SHAPE_LOG = CALL_LOG, None, shape_log_prev
    groups: dict = {}
    for name, cfn, args, shape in calls:
        groups.setdefault((name, shape), []).append((cfn, args))
    acc: dict = {}
    shapes = []
    for (name, shape), lst in groups.items():
        cfn, args = lst[0]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(reps):
                _lib.check(cfn(*args, st), name)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        a = acc.setdefault(name, {"ms": 0.0, "calls": 0})
        a["ms"] += ms * len(lst)
        a["calls"] += len(lst)
        shapes.append({"op": name, "shape": [str(v) for v in shape], "calls": len(lst), "us_per_call": 1e3 * ms})
    shapes.sort(key=lambda r: -r["us_per_call"] * r["calls"])
    del keep
    return {"by_op": acc, "by_shape": shapes, "total_ms": sum(a["ms"] for a in acc.values())}


def profile_step(runner, warm: int = 2, reps: int = 20) -> dict:
    """Live per-entry-point device time of ONE sampler step (udt_cfg_pack -> UNet -> udt_cfg_euler_step).

    The step is executed once eagerly while every C-ABI call (function, arguments) is recorded; each distinct
    (entry point, problem shape) is then replayed `reps` times back to back inside a CUDA graph on the launching
    stream and timed with CUDA events around the replay (no host launch cost, no per-call event overhead, programmatic
    dependent launch active exactly as in the product's step graph); a shape's time is multiplied by its call count.
    Also returns the duration of the whole step as a CUDA-graph replay.  The runner's state is restored afterwards."""
    x_saved = runner.x.clone()
    runner.row.copy_(runner.table[0:1])
    for _ in range(warm):
        runner._body()
    torch.cuda.synchronize()
    res = profile_callable(runner._body, reps)
    acc, shapes = res["by_op"], res["by_shape"]
    total = sum(a["ms"] for a in acc.values())
    step_ms_graph = None
    if runner.graph is not None:
        for _ in range(warm):
            runner.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            runner.replay()
        e1.record()
        torch.cuda.synchronize()
        step_ms_graph = e0.elapsed_time(e1) / 5
    runner.x.copy_(x_saved)
    return {"by_op": acc, "step_ms_eager_sum": total, "step_ms_graph": step_ms_graph, "by_shape": shapes}


def profile_step_in_graph(runner, replays: int = 12) -> dict:
    """Per-entry-point device time of one sampler step measured INSIDE a captured graph of the whole step: the step is
    captured once more with an external timing event (an event-record graph node) after every C-ABI call; the graph is
    replayed `replays` times back to back (clocks settle at what a long request sees) and the node-to-node intervals of
    the last replay are summed per entry point.  An interval spans from the completion of the previous call to the
    completion of this one, so it contains the launch gap the event node puts between two kernels (programmatic dependent
    launch cannot overlap across it): the sums are upper bounds of the kernels' time in the product's step graph, whose
    duration is returned beside them (`step_ms_graph_with_events` vs the plain graph)."""
    x_saved = runner.x.clone()
    runner.row.copy_(runner.table[0:1])
    runner._body()
    torch.cuda.synchronize()
    marks = []          # (entry point, event recorded after the call)

    g = torch.cuda.CUDAGraph()
    orig_invoke = globals()["_invoke"]

    def traced(name, *args, shape=None):
        orig_invoke(name, *args, shape=shape)
        e = torch.cuda.Event(enable_timing=True, external=True)
        e.record()
        marks.append((name, e))

    with torch.cuda.graph(g):
        start = torch.cuda.Event(enable_timing=True, external=True)
        start.record()
        globals()["_invoke"] = traced
        try:
            runner._body()
        finally:
            globals()["_invoke"] = orig_invoke
    for _ in range(replays):
        g.replay()
    torch.cuda.synchronize()
    acc: dict = {}
    prev = start
    for name, e in marks:
        a = acc.setdefault(name, {"ms": 0.0, "calls": 0})
        a["ms"] += prev.elapsed_time(e)
        a["calls"] += 1
        prev = e
    total = start.elapsed_time(marks[-1][1])
    runner.x.copy_(x_saved)
    return {"by_op": acc, "step_ms_graph_with_events": total}


_SPLITK_WS = {}   # device index -> fp32 scratch for split-K partial tiles (stream ordered, shared by all calls)
SPLITK_WS_BYTES = 64 << 20


def _splitk_ws(device: torch.device) -> torch.Tensor:
    key = device.index if device.index is not None else torch.cuda.current_device()
    ws = _SPLITK_WS.get(key)
    if ws is None:
        ws = torch.empty(SPLITK_WS_BYTES // 4, device=device, dtype=torch.float32)
        _SPLITK_WS[key] = ws
    return ws


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous CUDA {dtype} tensor, got {t.dtype} {t.device} "
                         f"contiguous={t.is_contiguous()}")


def igemm(
    srcs: Sequence[tuple],
    nb: int, h: int, w: int,
    weight: torch.Tensor,
    n_out: int,
    out: torch.Tensor,
    ldo: int,
    bias: Optional[torch.Tensor] = None,
    rowbias: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    ldr: int = 0,
    ld_rowbias: Optional[int] = None,
    out_fp32: bool = False,
    act: int = UDT_ACT_NONE,
    bn_hint: int = 0,
    out_ptr: Optional[int] = None,
    out_strides: Optional[Tuple[int, int, int]] = None,
    weight_img_rows: int = 0,
) -> torch.Tensor:
    """Segmented implicit GEMM (udt_igemm).  `srcs` = [(tensor, C, ld, taps[, stride, pad, H_in, W_in]), ...];
    (nb, h, w) are the OUTPUT pixel dims; 3x3 segments default to stride 1 / pad 1."""
    d = IGemmDesc()
    for i, src in enumerate(srcs):
        t, c, ld, taps = src[:4]
        g = d.src[i]
        g.ptr, g.C, g.ld, g.taps = t.data_ptr(), c, ld, taps
        g.stride = src[4] if len(src) > 4 else 1
        g.pad = src[5] if len(src) > 5 else (1 if taps == 9 else 0)
        g.H = src[6] if len(src) > 6 else 0
        g.W = src[7] if len(src) > 7 else 0
    d.nsrc, d.NB, d.H, d.W = len(srcs), nb, h, w
    d.weight, d.ldw, d.N_out = weight.data_ptr(), weight.stride(0), n_out
    d.bias, d.rowbias = _ptr(bias), _ptr(rowbias)
    d.ld_rowbias = 0 if rowbias is None else (n_out if ld_rowbias is None else ld_rowbias)
    d.residual, d.ldr = _ptr(residual), ldr
    d.out, d.ldo, d.out_fp32, d.act, d.bn_hint = out.data_ptr(), ldo, int(out_fp32), act, bn_hint
    d.weight_img_rows = weight_img_rows
    if out_strides is not None:       # strided output view (element strides of w, h, n), base pointer `out_ptr`
        d.out = out_ptr
        d.out_stride_w, d.out_stride_h, d.out_stride_n = out_strides
    ws = _splitk_ws(out.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    tuning = _igemm_tuning()
    if bn_hint == 0 and (tuning or KEY_LOG is not None):
        key = "|".join([f"{nb}x{h}x{w}", str(n_out), "+".join(f"{src[3]}x{src[1]}s{src[4] if len(src) > 4 else 1}" for src in srcs),
                        str(act), "r" if residual is not None else "-", "b" if rowbias is not None else "-",
                        "f32" if out_fp32 else "f16", "v" if out_strides is not None else "-", str(weight_img_rows)])
        if KEY_LOG is not None:
            KEY_LOG.append(key)
        d.bn_hint = tuning.get(key, 0)
    elif KEY_LOG is not None:
        KEY_LOG.append("fixed")
    if SHAPE_LOG is not None:
        k = sum(src[3] * ((src[1] + 63) // 64 * 64) for src in srcs)
        tag = "+".join(f"{src[3]}x{src[1]}" + (f"s{src[4]}" if len(src) > 4 and src[4] != 1 else "") for src in srcs)
        _invoke("udt_igemm", ctypes.byref(d), shape=(nb * h * w, n_out, k, tag, f"{nb}x{h}x{w}", act))
        return out
    _invoke("udt_igemm", ctypes.byref(d))
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, act: int = UDT_ACT_NONE,
           out_fp32: bool = False, bn_hint: int = 0, rowbias: Optional[torch.Tensor] = None, groups: int = 1,
           weight_img_rows: int = 0) -> torch.Tensor:
    """y[M, N] = act(x[M, K] @ weight[N, K]^T + bias) (+ residual); fp16 in, fp16 (or fp32) out.  `x` and `weight`
    may be row-strided 2-D views (unit column stride).  `rowbias` fp32 [G, N] adds row g to the g-th block of M / G
    consecutive rows (a per-sample bias).  `weight_img_rows` > 0: the g-th of `groups` row blocks multiplies with weight rows
    [g * weight_img_rows, g * weight_img_rows + N) (per-sample weights, N = weight_img_rows)."""
    m, k = x.shape
    n = weight_img_rows if weight_img_rows > 0 else weight.shape[0]
    n_log = n // 2 if act == UDT_ACT_GEGLU else n
    if act == UDT_ACT_GEGLU and bn_hint == 0:
        from .pack import GEGLU_TILE
        bn_hint = GEGLU_TILE          # the column interleave the weight was packed with
    if out is None:
        out = torch.empty((m, n_log), device=x.device, dtype=torch.float32 if out_fp32 else torch.float16)
    g = groups if rowbias is None else rowbias.shape[0]
    assert m % g == 0
    return igemm([(x, k, x.stride(0), 1)], g, 1, m // g, weight, n, out, out.stride(0), bias=bias, residual=residual,
                 ldr=0 if residual is None else residual.stride(0), out_fp32=out_fp32, act=act, bn_hint=bn_hint,
                 rowbias=rowbias, ld_rowbias=None if rowbias is None else rowbias.stride(0), weight_img_rows=weight_img_rows)


def conv3x3(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
            rowbias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
            skip_srcs: Sequence[torch.Tensor] = (), out: Optional[torch.Tensor] = None, out_fp32: bool = False,
            bn_hint: int = 0, stride: int = 1, pad: int = 1, ld_rowbias: Optional[int] = None) -> torch.Tensor:
    """3x3 conv on NHWC fp16 (stride 1 or 2; pad = low-side zero padding, the high side is padded as needed) with
    packed weight [Cout, 9*round_up(Cin,64) (+ skip segments)]; `skip_srcs` are extra NHWC tensors of the OUTPUT
    size consumed point-wise (fused 1x1 skip connection)."""
    nb, h, w, c = x.shape
    ho, wo = ((h + 2 * pad - 3) // stride + 1, (w + 2 * pad - 3) // stride + 1) if pad else (h // stride, w // stride)
    n = weight.shape[0]
    if out is None:
        out = torch.empty((nb, ho, wo, n), device=x.device, dtype=torch.float32 if out_fp32 else torch.float16)
    srcs = [(x, c, x.stride(2), 9, stride, pad, h, w)] + [(s, s.shape[-1], s.stride(2), 1) for s in skip_srcs]
    if rowbias is not None and ld_rowbias is None:
        ld_rowbias = rowbias.stride(0) if rowbias.dim() == 2 else 0
    return igemm(srcs, nb, ho, wo, weight, n, out, out.stride(2), bias=bias, rowbias=rowbias, residual=residual,
                 ldr=0 if residual is None else residual.stride(2), out_fp32=out_fp32, bn_hint=bn_hint,
                 ld_rowbias=ld_rowbias)


def conv3x3_up2(x: torch.Tensor, weights4: Sequence[torch.Tensor], bias: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nearest-2x upsample + 3x3 conv (pad 1) of NHWC fp16 `x` [nb, h, w, c] -> [nb, 2h, 2w, n] as four 2x2-window implicit
    GEMMs on the low-resolution input, one per output phase (`weights4` from pack.pack_conv3x3_up2): 2.25x fewer FLOPs
    than materialising the upsampled tensor, and no upsample kernel."""
    nb, h, w, c = x.shape
    n = weights4[0].shape[0]
    if out is None:
        out = torch.empty((nb, 2 * h, 2 * w, n), device=x.device, dtype=torch.float16)
    sw, sh, sn = 2 * n, 2 * (2 * w) * n, (2 * h) * (2 * w) * n
    for ph, wt in enumerate(weights4):
        py, px = ph >> 1, ph & 1
        pad = 2 * (1 - py) + (1 - px)            # window rows {h-1, h} for py = 0, {h, h+1} for py = 1 (same in x)
        base = out.data_ptr() + 2 * (py * (2 * w) + px) * n
        igemm([(x, c, x.stride(2), 4, 1, pad, h, w)], nb, h, w, wt, n, out, n, bias=bias, out_ptr=base, out_strides=(sw, sh, sn))
    return out


def groupnorm_ws_bytes(nb: int, hw: int, c: int, groups: int = 32) -> int:
    return int(_lib.load().udt_groupnorm_ws_bytes(nb, hw, c, groups))


def groupnorm(x0: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, silu: bool,
              x1: Optional[torch.Tensor] = None, groups: int = 32, out: Optional[torch.Tensor] = None,
              ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm(+SiLU) over NHWC fp16; with `x1` normalises cat([x0, x1], channel) and writes the concat."""
    nb = x0.shape[0]
    c0 = x0.shape[-1]
    c1 = 0 if x1 is None else x1.shape[-1]
    hw = x0.numel() // (nb * c0)
    if out is None:
        out = torch.empty(tuple(x0.shape[:-1]) + (c0 + c1,), device=x0.device, dtype=torch.float16)
    if ws is None or ws.numel() * 8 < groupnorm_ws_bytes(nb, hw, c0 + c1, groups):
        ws = torch.empty(groupnorm_ws_bytes(nb, hw, c0 + c1, groups) // 8, device=x0.device, dtype=torch.float64)
    _invoke("udt_groupnorm_nhwc", x0.data_ptr(), c0, _ptr(x1), c1, out.data_ptr(), nb, hw, groups, gamma.data_ptr(),
                              beta.data_ptr(), float(eps), int(silu), ws.data_ptr())
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    c = x.shape[-1]
    rows = x.numel() // c
    if out is None:
        out = torch.empty_like(x)
    _invoke("udt_layernorm", x.data_ptr(), out.data_ptr(), rows, c, gamma.data_ptr(), beta.data_ptr(), float(eps))
    return out


def fmha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, b: int, nq: int, nkv: int, heads: int, scale: float,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q/k/v: 2-D fp16 views [B*N, >= heads*64] (may be column slices of one fused QKV buffer)."""
    if out is None:
        out = torch.empty((b * nq, heads * 64), device=q.device, dtype=torch.float16)
    _invoke("udt_fmha_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), b, nq, nkv, heads, q.stride(0),
                        k.stride(0), v.stride(0), out.stride(0), float(scale))
    return out


def xattn_small_l(q: torch.Tensor, kc: torch.Tensor, vc: torch.Tensor, b: int, n: int, l: int, heads: int,
                  scale: float, out: Optional[torch.Tensor] = None, probs: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B*N, heads*64]; kc/vc [B*L, heads*64] (row pitch may be larger); probs fp32 [B*heads, N, L] optional."""
    if out is None:
        out = torch.empty((b * n, heads * 64), device=q.device, dtype=torch.float16)
    _invoke("udt_xattn_small_l", q.data_ptr(), kc.data_ptr(), vc.data_ptr(), out.data_ptr(), _ptr(probs), b, n, l, heads,
                             q.stride(0), kc.stride(0), out.stride(0), float(scale))
    return out


def xattn_fold(kc: torch.Tensor, vc: torch.Tensor, wq: torch.Tensor, wo: torch.Tensor, b: int, l: int, heads: int,
               scale: float, w1: torch.Tensor, w2: torch.Tensor) -> None:
    """fold the step-invariant context K / V (fp16 views [b*l, heads*64]) into the t_attn projections:
    w1 fp16 [b, npad, C] (scores = LN(t) @ w1^T), w2 fp16 [b, C, npad] (out = P @ w2^T); see include/udt_api.h"""
    npad = w1.shape[1]
    _invoke("udt_xattn_fold", kc.data_ptr(), vc.data_ptr(), kc.stride(0), wq.data_ptr(), wq.stride(0), wo.data_ptr(),
            wo.stride(0), w1.data_ptr(), w2.data_ptr(), b, l, heads, npad, float(scale))


def softmax_groups(x: torch.Tensor, groups: int, l: int, out: Optional[torch.Tensor] = None,
                   probs: Optional[torch.Tensor] = None, n: int = 0) -> torch.Tensor:
    """x fp16 [rows, cols]: softmax over `groups` groups of `l` consecutive columns, pad columns zeroed (in place by default)"""
    rows, cols = x.shape
    if out is None:
        out = x
    _invoke("udt_softmax_groups", x.data_ptr(), out.data_ptr(), rows, cols, x.stride(0), groups, l, _ptr(probs), n)
    return out


def request_pack_u8(image_u8: torch.Tensor, mask_u8: torch.Tensor, b: int):
    """uint8 [Bs, H, W, 3] image + uint8 [Bs, H, W, MC] user mask (device) -> (image, mask, masked) fp32 NCHW, demo.py:52-62"""
    bs, hh, ww, _ = image_u8.shape
    assert image_u8.dtype == mask_u8.dtype == torch.uint8 and image_u8.is_contiguous() and mask_u8.is_contiguous()
    assert image_u8.shape[3] == 3 and mask_u8.shape[:3] == image_u8.shape[:3]
    dev = image_u8.device
    image = torch.empty((b, 3, hh, ww), device=dev, dtype=torch.float32)
    masked = torch.empty_like(image)
    mask = torch.empty((b, 1, hh, ww), device=dev, dtype=torch.float32)
    _invoke("udt_request_pack_u8", image_u8.data_ptr(), mask_u8.data_ptr(), image.data_ptr(), mask.data_ptr(), masked.data_ptr(),
            b, bs, hh, ww, mask_u8.shape[3])
    return image, mask, masked


def images_to_u8(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW [B, C, H, W] in [0, 1] -> uint8 NHWC, trunc(x * 255) (demo.py:100-101)"""
    b, c, hh, ww = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty((b, hh, ww, c), device=x.device, dtype=torch.uint8)
    _invoke("udt_images_to_u8", x.data_ptr(), y.data_ptr(), b, c, hh * ww)
    return y


def attn_local_score(probs: torch.Tensor, mask: torch.Tensor, seg: torch.Tensor, gk: torch.Tensor, score: torch.Tensor,
                     heads: int, size: int) -> None:
    """score[b] += noise-search score of one t_attn layer (K12, loss.py:192-235): probs fp32 [B*heads, size*size, L],
    mask fp32 [Bm, 1, H, W], seg fp32 [Bm, seg_l], gk fp32 [ks, ks], score fp32 [B] (B = Bm or 2*Bm)"""
    b = score.shape[0]
    bm, _, hh, ww = mask.shape
    assert probs.is_contiguous() and mask.is_contiguous() and seg.is_contiguous() and gk.is_contiguous()
    assert probs.dtype == mask.dtype == seg.dtype == gk.dtype == score.dtype == torch.float32
    assert probs.shape[0] == b * heads and probs.shape[1] == size * size
    _invoke("udt_attn_local_score", probs.data_ptr(), mask.data_ptr(), seg.data_ptr(), gk.data_ptr(), score.data_ptr(),
            b, bm, heads, size, probs.shape[2], seg.shape[1], hh, ww, gk.shape[-1])


def label_embed(idx: torch.Tensor, emb: torch.Tensor, pe: torch.Tensor, out: Optional[torch.Tensor] = None,
                out_f32: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None) -> torch.Tensor:
    """idx int32 [B, L]; emb fp32 [V, D]; pe fp32 [L, D] -> fp16 [B*L, D] (optionally also fp32 and the low fp16 half)"""
    b, l = idx.shape
    d = emb.shape[1]
    if out is None:
        out = torch.empty((b * l, d), device=emb.device, dtype=torch.float16)
    _invoke("udt_label_embed", idx.data_ptr(), emb.data_ptr(), pe.data_ptr(), out.data_ptr(), _ptr(out_f32), _ptr(out_lo),
            b * l, l, d)
    return out


def rowsum_norm_split(parts: Sequence[torch.Tensor], res: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
                      beta: Optional[torch.Tensor] = None, eps: float = 1e-5, relu: bool = False,
                      out_f32: Optional[torch.Tensor] = None, out_hi: Optional[torch.Tensor] = None,
                      out_lo: Optional[torch.Tensor] = None) -> None:
    """y = LN(relu?(sum(parts)) + res) row-wise over fp32 [rows, C] tensors (1-3 parts; LN optional) -> fp32 and / or the
    fp16 pair hi / lo (udt_rowsum_norm_split)"""
    rows, c = parts[0].shape
    for t in list(parts) + [t for t in (res, out_f32) if t is not None]:
        assert t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (rows, c)
    for t in (out_hi, out_lo):
        assert t is None or (t.dtype == torch.float16 and t.is_contiguous() and tuple(t.shape) == (rows, c))
    p = [t.data_ptr() for t in parts] + [None] * (3 - len(parts))
    _invoke("udt_rowsum_norm_split", p[0], p[1], p[2], _ptr(res), rows, c, _ptr(gamma), _ptr(beta), float(eps), int(relu),
            _ptr(out_f32), _ptr(out_hi), _ptr(out_lo))


def mha_small_f32(qkv: torch.Tensor, b: int, l: int, heads: int, out_hi: torch.Tensor, out_lo: torch.Tensor) -> None:
    """qkv fp32 [B*L, 3*D] (q | k | v) -> the fp16 pair out_hi / out_lo [B*L, D]; softmax scale = head_dim^-0.5"""
    d = qkv.shape[1] // 3
    dh = d // heads
    assert qkv.dtype == torch.float32 and qkv.is_contiguous()
    _invoke("udt_mha_small_f32", qkv.data_ptr(), out_hi.data_ptr(), out_lo.data_ptr(), b, l, heads, dh, qkv.stride(0),
            out_hi.stride(0), float(dh) ** -0.5)


def mha_small(qkv: torch.Tensor, b: int, l: int, heads: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv fp16 [B*L, 3*D] (q | k | v) -> fp16 [B*L, D]; softmax scale = head_dim^-0.5"""
    d = qkv.shape[1] // 3
    dh = d // heads
    if out is None:
        out = torch.empty((b * l, d), device=qkv.device, dtype=torch.float16)
    _invoke("udt_mha_small", qkv.data_ptr(), out.data_ptr(), b, l, heads, dh, qkv.stride(0), out.stride(0),
                               float(dh) ** -0.5)
    return out


def mha_masked(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, b: int, lq: int, lk: int, heads: int,
               mask: Optional[torch.Tensor] = None, kpm: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q fp16 [B*Lq, >= D], k / v fp16 [B*Lk, >= D] (2-D views, unit column stride) -> fp16 [B*Lq, D]; mask fp32 [Lq, Lk]
    additive, kpm uint8 [B, Lk] (udt_mha_masked; softmax scale = head_dim^-0.5)"""
    d = q.shape[1]
    dh = d // heads
    if out is None:
        out = torch.empty((b * lq, d), device=q.device, dtype=torch.float16)
    assert mask is None or (mask.dtype == torch.float32 and mask.stride(1) == 1 and tuple(mask.shape) == (lq, lk))
    assert kpm is None or (kpm.dtype == torch.uint8 and kpm.is_contiguous() and tuple(kpm.shape) == (b, lk))
    _invoke("udt_mha_masked", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), b, lq, lk, heads, dh, q.stride(0),
            k.stride(0), v.stride(0), out.stride(0), float(dh) ** -0.5, _ptr(mask), 0 if mask is None else mask.stride(0),
            _ptr(kpm))
    return out


def softmax_rows_(x: torch.Tensor, scale: float) -> torch.Tensor:
    rows, cols = x.shape
    _invoke("udt_softmax_rows", x.data_ptr(), rows, cols, x.stride(0), float(scale))
    return x


def cfg_pack(x: torch.Tensor, cat_uc: torch.Tensor, cat_c: torch.Tensor, c_in_dev: torch.Tensor, out: torch.Tensor
             ) -> torch.Tensor:
    """x fp32 NCHW [B,4,h,w], cat_* fp32 NCHW [B,5,h,w], c_in_dev: device fp32 scalar -> out fp16 NHWC [2B,h,w,16]"""
    b = x.shape[0]
    hw = x.numel() // (b * 4)
    _invoke("udt_cfg_pack", x.data_ptr(), cat_uc.data_ptr(), cat_c.data_ptr(), out.data_ptr(), b, hw,
                              c_in_dev.data_ptr())
    return out


def cfg_euler_step_(x: torch.Tensor, eps2b: torch.Tensor, cfg_scale: float, dsigma_dev: torch.Tensor,
                    cfg_scale_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x fp32 NCHW [B,4,h,w] += dsigma * cfg(eps2b fp32 NHWC [2B,h,w,4]); dsigma_dev: device fp32 scalar; `cfg_scale_dev`
    (device fp32 scalar) overrides `cfg_scale` when given"""
    b = x.shape[0]
    hw = x.numel() // (b * 4)
    _invoke("udt_cfg_euler_step", x.data_ptr(), eps2b.data_ptr(), b, hw, float(cfg_scale), dsigma_dev.data_ptr(),
            _ptr(cfg_scale_dev))
    return x


def vae_sample_pack(moments: torch.Tensor, noise_c: torch.Tensor, noise_uc: torch.Tensor, mask: torch.Tensor,
                    scale_factor: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """moments fp32 NHWC [B,h,w,>=8]; noise fp32 NCHW [B,4,h,w]; mask fp32 [B,1,8h,8w] -> (concat_c, concat_uc) NCHW"""
    b, h, w, ld = moments.shape
    cat_c = torch.empty((b, 5, h, w), device=moments.device, dtype=torch.float32)
    cat_uc = torch.empty_like(cat_c)
    _invoke("udt_vae_sample_pack", moments.data_ptr(), ld, noise_c.data_ptr(), noise_uc.data_ptr(), mask.data_ptr(),
                                     cat_c.data_ptr(), cat_uc.data_ptr(), b, h, w, float(scale_factor))
    return cat_c, cat_uc


def pointwise_affine(x: torch.Tensor, wm: torch.Tensor, bias: Optional[torch.Tensor], cpad: int, in_scale: float
                     ) -> torch.Tensor:
    """x fp32 NCHW [B,Cin,h,w] -> fp16 NHWC [B,h,w,cpad] = wm @ (x * in_scale) + bias"""
    b, cin, h, w = x.shape
    out = torch.empty((b, h, w, cpad), device=x.device, dtype=torch.float16)
    _invoke("udt_pointwise_affine", x.data_ptr(), wm.data_ptr(), _ptr(bias), out.data_ptr(), b, h * w, cin, wm.shape[0],
                                      cpad, float(in_scale))
    return out


def upsample2x(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    nb, h, w, c = x.shape
    if out is None:
        out = torch.empty((nb, 2 * h, 2 * w, c), device=x.device, dtype=torch.float16)
    _invoke("udt_upsample2x_nhwc", x.data_ptr(), out.data_ptr(), nb, h, w, c)
    return out


def nchw_to_nhwc_f16(x: torch.Tensor, cpad: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    nb, c, h, w = x.shape
    cpad = c if cpad is None else cpad
    if out is None:
        out = torch.empty((nb, h, w, cpad), device=x.device, dtype=torch.float16)
    _invoke("udt_nchw_f32_to_nhwc_f16", x.data_ptr(), out.data_ptr(), nb, c, h * w, cpad)
    return out


def nhwc_to_nchw_f32(x: torch.Tensor, c: int, scale: float = 1.0, shift: float = 0.0, clamp01: bool = False,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    nb, h, w, ld = x.shape
    if out is None:
        out = torch.empty((nb, c, h, w), device=x.device, dtype=torch.float32)
    _invoke("udt_nhwc_to_nchw_f32", x.data_ptr(), int(x.dtype == torch.float32), out.data_ptr(), nb, c, h * w, ld,
                                      float(scale), float(shift), int(clamp01))
    return out
