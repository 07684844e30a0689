"""ctypes binding of libot_b200.so (declared in include/ot_b200.h).

The library is loaded from the package directory; if it is missing, loading raises -- there is no eager /
PyTorch / CPU fallback for any op (north-star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# OT_B200_LIB: a development override (A/B builds of one kernel, tools/build_variant.py); the library must exist either way
LIB_PATH = os.environ.get("OT_B200_LIB") or os.path.join(HERE, "libot_b200.so")

OT_OK, OT_EINVAL, OT_ECUDA, OT_ENODEV = 0, -1, -2, -3

FAULT_NONE, FAULT_INPUT, FAULT_WEIGHT, FAULT_RANDOM_BITFLIP, FAULT_RANDOM, FAULT_ACC_BITFLIP, FAULT_OUT_Q8_BITFLIP = range(7)
OUT_I32, OUT_F32, OUT_Q8, OUT_QLINEAR = 0, 1, 2, 3


class OtFault(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("bit", C.c_int32),
        ("flat_index", C.c_int64),
        ("window_start", C.c_int32),
        ("window_len", C.c_int32),
        ("value_bits", C.c_uint32),
        ("reserved", C.c_int32),
    ]


class OtError(RuntimeError):
    pass


_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float
_I4 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol include/ot_b200.h declares is listed (tests check the export table).
SIGNATURES = {
    "ot_version": (_i, []),
    "ot_last_error": (C.c_char_p, []),
    "ot_device_ok": (_i, []),
    "ot_launch_count": (_l, []),
    "ot_set_pdl": (_i, [_i]),
    "ot_set_timeline": (_i, [_p, C.c_uint]),
    "ot_linear_w8a8": (_i, [_p, _l, _p, _l, _i, _i, _i, _p, _p, _p, _p, _l, _i, _i, _p, _l, _p, _i, C.POINTER(OtFault), _p]),
    "ot_linear_w8a8_mf": (_i, [_p, _l, _p, _l, _i, _i, _i, _p, _p, _p, _p, _l, _i, _i, _p, _l, _p, _i, _p, _p, _i, _p]),
    "ot_linear_w4a8": (_i, [_p, _l, _p, _l, _i, _i, _i, _p, _p, _p, _p, _l, _i, _i, _p, _l, _p, _i, C.POINTER(OtFault), _p]),
    "ot_linear_w4a8_mf": (_i, [_p, _l, _p, _l, _i, _i, _i, _p, _p, _p, _p, _l, _i, _i, _p, _l, _p, _i, _p, _p, _i, _p]),
    "ot_matmul_integer": (_i, [_p, _l, _p, _l, _i, _i, _i, _p, _p, _p, _p, _p, _l, C.POINTER(OtFault), _p]),
    "ot_qlinear_matmul": (_i, [_p, _l, _p, _l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _f, _i, _p, _l, _p]),
    "ot_rowsum_i8": (_i, [_p, _l, _l, _i, _p, _p]),
    "ot_ln_linear_w8a8": (_i, [_p, _l, _p, _p, _f, _p, _l, _i, _i, _i, _p, _p, _p, _l, _i, _i, _p, _l, _p, _i, _p]),
    "ot_unpack_int4": (_i, [_p, _p, _l, _l, _p]),
    "ot_pack_int4": (_i, [_p, _p, _l, _l, _p]),
    "ot_layernorm_quant": (_i, [_p, _p, _p, _l, _i, _f, _p, _p, _p, _p]),
    "ot_rowquant": (_i, [_p, _l, _l, _i, _i, _p, _p, _p, _p]),
    "ot_residual_add": (_i, [_p, _p, _p, _l, _p]),
    "ot_embed_pe": (_i, [_p, _l, _p, _p, _l, _i, _i, _i, _p, _f, _p, _p]),
    "ot_attention_q8": (_i, [_p, _l, _p, _l, _p, _p, _l, _p, _p, _l, _p, _p, _l, _p, _p, _l,
                               _i, _i, _i, _i, _i, _i, _p, _l, _i, _p, _p, _l, _p, _p, _p, C.POINTER(OtFault), _p]),
    "ot_attention_q8_mf": (_i, [_p, _l, _p, _l, _p, _p, _l, _p, _p, _l, _p, _p, _l, _p, _p, _l,
                                  _i, _i, _i, _i, _i, _i, _p, _l, _i, _p, _p, _l, _p, _p, _p, C.POINTER(OtFault), _p, _p, _p]),
    "ot_generator_argmax": (_i, [_p, _l, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ot_append_token": (_i, [_p, _l, _p, _i, _p, _p]),
    "ot_decoder_plan_size": (_i, []),
    "ot_decoder_plan_build": (_i, [_p, _i, _i, _i, _i, _i, _l, C.POINTER(_p), C.POINTER(_p)]),
    "ot_decoder_run": (_i, [_p, _p, _i, _i, _p]),
    "ot_cdecoder_plan_size": (_i, []),
    "ot_cdecoder_plan_build": (_i, [_p, _i, _i, _i, _i, _i, _i, _l, C.POINTER(_p), C.POINTER(_p)]),
    "ot_cdecoder_run": (_i, [_p, _i, _i, _i, _i, _p]),
    "ot_cdecoder_max_clusters": (_i, [C.POINTER(C.c_int)]),
    "ot_unary_f32": (_i, [_i, _p, _p, _l, _p]),
    "ot_binary_f32": (_i, [_i, _p, _I4, _p, _I4, _p, _I4, _p]),
    "ot_clip_f32": (_i, [_p, _f, _f, _p, _l, _p]),
    "ot_reduce_last_f32": (_i, [_i, _p, _l, _i, _p, _p]),
    "ot_softmax_f32": (_i, [_p, _l, _i, _p, _p]),
    "ot_where_f32": (_i, [_p, _I4, _f, _p, _p, _I4, _p]),
    "ot_equal_i64": (_i, [_p, _l, _p, _l, _p]),
    "ot_cast": (_i, [_i, _p, _i, _p, _l, _p]),
    "ot_transpose4_b32": (_i, [_p, _I4, C.POINTER(C.c_int), _p, _p]),
    "ot_matmul_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _l, _l, _l, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and bind every declared symbol. Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OtError(
            "libot_b200.so is not built (%s). Run `python __graft_entry__.py build` -- this package has no "
            "CPU or eager fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != OT_OK:
        msg = load().ot_last_error().decode(errors="replace")
        raise OtError("%s failed (rc=%d): %s" % (what, rc, msg))


def launch_count() -> int:
    return int(load().ot_launch_count())
