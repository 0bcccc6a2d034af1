"""ctypes binding of libudt_b200.so (include/udt_api.h).

The library is the only compute path of this package: if it is missing or the device is not sm_100 every call
raises — there is no PyTorch / CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libudt_b200.so")

UDT_ACT_NONE, UDT_ACT_SILU, UDT_ACT_GEGLU, UDT_ACT_RELU, UDT_ACT_GELU = 0, 1, 2, 3, 4


class UdtError(RuntimeError):
    pass


class GemmSrc(Structure):
    """mirror of `udt_gemm_src`"""

    _fields_ = [("ptr", c_void_p), ("C", c_int32), ("ld", c_int32), ("taps", c_int32), ("H", c_int32), ("W", c_int32),
                ("stride", c_int32), ("pad", c_int32)]


class IGemmDesc(Structure):
    """mirror of `udt_igemm_desc`"""

    _fields_ = [("src", GemmSrc * 3), ("nsrc", c_int32), ("NB", c_int32), ("H", c_int32), ("W", c_int32),
                ("weight", c_void_p), ("ldw", c_int32), ("N_out", c_int32), ("bias", c_void_p), ("rowbias", c_void_p),
                ("ld_rowbias", c_int32), ("residual", c_void_p), ("ldr", c_int32), ("out", c_void_p), ("ldo", c_int32),
                ("out_fp32", c_int32), ("act", c_int32), ("bn_hint", c_int32), ("out_stride_w", c_int64), ("out_stride_h", c_int64),
                ("out_stride_n", c_int64), ("weight_img_rows", c_int32), ("workspace", c_void_p),
                ("workspace_bytes", c_int64)]


# name -> (restype, argtypes): every symbol include/udt_api.h declares
_PROTOTYPES = {
    "udt_version": (c_int32, []),
    "udt_arch": (c_int32, []),
    "udt_last_error": (c_char_p, []),
    "udt_num_sms": (c_int32, []),
    "udt_sizeof_igemm_desc": (c_int32, []),
    "udt_geglu_tile": (c_int32, []),
    "udt_igemm": (c_int32, [POINTER(IGemmDesc), c_void_p]),
    "udt_debug_set_trace": (c_int32, [c_void_p, c_int64]),
    "udt_groupnorm_ws_bytes": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "udt_groupnorm_nhwc": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                                     c_void_p, c_float, c_int32, c_void_p, c_void_p]),
    "udt_layernorm": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_float, c_void_p]),
    "udt_fmha_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                               c_int32, c_int32, c_float, c_void_p]),
    "udt_xattn_small_l": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                    c_int32, c_int32, c_int32, c_float, c_void_p]),
    "udt_request_pack_u8": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                      c_int32, c_void_p]),
    "udt_images_to_u8": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "udt_attn_local_score": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                       c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "udt_label_embed": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "udt_rowsum_norm_split": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_float,
                                        c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "udt_mha_small_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float,
                                    c_void_p]),
    "udt_mha_small": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "udt_mha_masked": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                 c_int32, c_int32, c_int32, c_float, c_void_p, c_int32, c_void_p, c_void_p]),
    "udt_softmax_rows": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "udt_xattn_fold": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_int32,
                                 c_int32, c_int32, c_int32, c_float, c_void_p]),
    "udt_softmax_groups": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "udt_cfg_pack": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "udt_cfg_euler_step": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p]),
    "udt_vae_sample_pack": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                      c_int32, c_float, c_void_p]),
    "udt_pointwise_affine": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                       c_float, c_void_p]),
    "udt_upsample2x_nhwc": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "udt_nchw_f32_to_nhwc_f16": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "udt_nhwc_to_nchw_f32": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_float,
                                       c_int32, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def load() -> ctypes.CDLL:
    """dlopen the in-tree library (built by `udifftext_b200.build`) and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UdtError(
            f"{LIB_PATH} not found: run `python -m udifftext_b200.build` (or __graft_entry__.build()); "
            "udifftext_b200 has no fallback compute path"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().udt_last_error()
        raise UdtError(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")
