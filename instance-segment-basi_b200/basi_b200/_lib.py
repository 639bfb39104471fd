"""ctypes binding of libbasi_b200.so (the C ABI declared in include/basi_b200.h).

There is no CPU fallback: every call goes to the CUDA library, and a missing
library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BASI_LIB") or os.path.join(_HERE, "libbasi_b200.so")   # BASI_LIB: A/B another build
# the same kernels with IEEE fp16 as the 16-bit storage type (csrc/build.sh, -DBASI_HALF_FP16): precision "f16"
LIB_PATHS = {"bf16": LIB_PATH, "f16": os.path.join(_HERE, "libbasi_b200_f16.so")}

F32, BF16 = 0, 1
TC_FPROP, TC_DGRAD, TC_WGRAD = 0, 1, 2


class BasiError(RuntimeError):
    pass


class Tensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("ld", C.c_int32), ("dtype", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("dil", C.c_int32),
                ("pad_t", C.c_int32), ("pad_l", C.c_int32), ("relu", C.c_int32)]


class PackEntry(C.Structure):
    _fields_ = [("w", C.c_void_p), ("w_io", C.c_void_p), ("w_oi", C.c_void_p), ("taps", C.c_int32),
                ("cin", C.c_int32), ("cout", C.c_int32), ("block_start", C.c_int32), ("tiles_co", C.c_int32),
                ("tiles_ci", C.c_int32), ("pad0", C.c_int32), ("pad1", C.c_int32)]


_P = C.c_void_p
_TP = C.POINTER(Tensor)
_DP = C.POINTER(ConvDesc)
_i, _i64, _f, _d = C.c_int, C.c_int64, C.c_float, C.c_double

# name -> argtypes.  Every symbol include/basi_b200.h declares is listed here; tests check the two agree.
SIGNATURES = {
    "basi_memset": [_P, _i, _i64, _P],
    "basi_label_encode": [_P, _P, _P, _i, _P, _P, _i, _i64, _P],
    "basi_click_count": [_P, _i, _i, _i, _i, _P, _P],
    "basi_click_select": [_P, _i, _i, _i, _i, _i, _P, _i, _P, _P],
    "basi_clickmap_pack": [_P, _i, _P, _P, _i64, _P, _i, _i, _i, _P],
    "basi_conv_fprop": [_DP, _TP, _P, _P, _TP, _P],
    "basi_conv_dgrad": [_DP, _TP, _P, _TP, _i, _P],
    "basi_conv_wgrad": [_DP, _TP, _TP, _P, _P, _P],
    "basi_bn_maskbits_supported": [_TP],
    "basi_bn_apply_bits": [_TP, _P, _TP, _P, _i, _TP, _P, _P],
    "basi_bn_bwd_reduce_bits": [_TP, _P, _TP, _P, _P, _d, _P, _P, _P, _P, _P],
    "basi_bn_bwd_apply_bits": [_TP, _P, _TP, _P, _P, _TP, _TP, _i, _P],
    "basi_bn_bwd_coop_supported": [_TP, _i, _i, _i],
    "basi_bn_bwd_coop": [_TP, _TP, _P, _TP, _P, _i, _P, _d, _P, _P, _P, _P, _TP, _TP, _i, _P],
    "basi_bn_bwd_fused_supported": [_TP],
    "basi_bn_bwd_fused": [_TP, _TP, _P, _i, _P, _d, _P, _P, _P, _P, _TP, _P],
    "basi_avgpool_multi_fwd": [_TP, _i, _P, _P, _P, _P],
    "basi_avgpool_multi_bwd": [_P, _i, _P, _TP, _i, _P],
    "basi_stem_fprop_stats_supported": [_DP, _TP, _TP],
    "basi_stem_fprop_stats": [_DP, _TP, _P, _TP, _P, _P, _P, _d, _f, _P, _P, _P],
    "basi_subsample_fwd": [_TP, _i, _TP, _P],
    "basi_subsample_bwd": [_TP, _i, _TP, _i, _P],
    "basi_bn_stats": [_TP, _P, _P, _P, _d, _f, _P, _P, _P],
    "basi_bn_finalize": [_P, _P, _P, _d, _f, _P, _i, _P],
    "basi_bn_apply": [_TP, _P, _TP, _P, _i, _TP, _P],
    "basi_bn_bwd_reduce": [_TP, _TP, _TP, _P, _i, _P, _d, _P, _P, _P, _P, _P],
    "basi_bn_bwd_finalize": [_P, _d, _P, _P, _P, _i, _P],
    "basi_bn_bwd_apply": [_TP, _TP, _TP, _P, _P, _i, _TP, _TP, _i, _P],
    "basi_bias_relu_bwd": [_TP, _TP, _i, _P, _P],
    "basi_tc_conv_set_bias": [_P, _P, _i],
    "basi_relu_fwd": [_TP, _TP, _P],
    "basi_relu_bwd": [_TP, _TP, _TP, _i, _P],
    "basi_add_fwd": [_TP, _TP, _TP, _P],
    "basi_add_bwd": [_TP, _TP, _i, _TP, _i, _P],
    "basi_onehot2_f32": [_P, _P, _i64, _P],
    "basi_maxpool2s2_fwd": [_TP, _TP, _P, _P],
    "basi_maxpool2s2_bwd": [_TP, _P, _TP, _i, _P],
    "basi_maxpool3s2_fwd": [_TP, _TP, _P, _P],
    "basi_maxpool3s2_bwd": [_TP, _P, _TP, _i, _P],
    "basi_avgpool_fwd": [_TP, _i, _TP, _P],
    "basi_avgpool_bwd": [_TP, _i, _TP, _i, _P],
    "basi_bilinear_ac_fwd": [_TP, _TP, _P],
    "basi_bilinear_ac_bwd": [_TP, _TP, _i, _P],
    "basi_gate_mul_fwd": [_TP, _P, _i, _i, _TP, _P],
    "basi_gate_mul_bwd": [_TP, _TP, _P, _i, _i, _TP, _i, _P, _P],
    "basi_mask_mul_fwd": [_TP, _P, _i, _i, _TP, _P],
    "basi_mask_mul_bwd": [_TP, _TP, _P, _i, _i, _TP, _i, _P, _P],
    "basi_softmax_gate_fwd": [_P, _i64, _i, _i, _f, _P, _P],
    "basi_softmax_gate_bwd": [_P, _P, _i64, _i, _i, _f, _P, _i, _P],
    "basi_resize_nearest_fwd": [_TP, _TP, _P],
    "basi_resize_nearest_bwd": [_TP, _TP, _i, _P],
    "basi_skinny_supported": [_i, _i, _i],
    "basi_skinny_fwd": [_P, _i, _i64, _P, _P, _P, _i, _i, _i, _i, _P],
    "basi_skinny_fwd_ws": [_P, _i, _i64, _P, _P, _P, _i, _i, _i, _i, _P, _P],
    "basi_skinny_dgrad": [_P, _P, _P, _i, _i64, _i, _i, _i, _i, _P],
    "basi_skinny_wgrad": [_P, _i, _i64, _P, _P, _P, _i, _i, _i, _P],
    "basi_relu_bwd_f32": [_P, _P, _i64, _P],
    "basi_wbce_fwd_bwd": [_P, _P, _f, _d, _f, _i64, _P, _P, _P],
    "basi_wbce_sel_fwd_bwd": [_P, _i, _i, _P, _f, _d, _f, _i64, _P, _P, _P],
    "basi_sigmoid_fwd": [_P, _P, _i64, _P],
    "basi_sigmoid_bwd": [_P, _P, _P, _i64, _i, _P],
    "basi_softmax_ce_fwd_bwd": [_P, _P, _i64, _i, _d, _f, _P, _P, _P],
    "basi_p2p_allreduce_mean": [_P, _P, _i, _i, _i64, _P, _P],
    "basi_sgd_step": [_P, _P, _P, _i64, _P, _P],
    "basi_threshold": [_P, _f, _P, _i64, _P],
    "basi_argmax": [_P, _i64, _i, _P, _P],
    "basi_upsample_legacy_argmax": [_P, _i, _i, _i, _i, _i, _i, _P, _P],
    "basi_split3_bf16": [_TP, _TP, _P],
    "basi_cast_f32_to_bf16": [_P, _P, _i64, _P],
    "basi_cast_bf16_to_f32": [_P, _P, _i64, _P],
    "basi_tc_conv_supported": [_i, _DP, _TP, _TP],
    "basi_tc_pack_weights": [_P, _P, _P, _i, _i, _i, _P],
    "basi_tc_pack_weights_multi": [_P, _i, _i, _P],
    "basi_tc_conv_create": [_i, _DP, _TP, _TP, _P, _P, _i, C.POINTER(_P)],
    "basi_tc_conv_supported_split": [_i, _DP, _TP, _TP],
    "basi_tc_conv_create_split": [_i, _DP, _TP, _TP, _P, _P, _i, _i, C.POINTER(_P)],
    "basi_tc_conv_set_bn_stats": [_P, _P, _P, _P, _d, _f, _P, _P],
    "basi_tc_conv_run": [_P, _P],
    "basi_set_sm_budget": [_i, _i],
}
_NOCHECK = {"basi_last_error": ([], C.c_char_p), "basi_version": ([], _i), "basi_sm_count": ([], _i),
            "basi_half_format": ([], _i),
            "basi_tc_conv_destroy": ([_P], None), "basi_tc_conv_set_bn_apply": ([_P, _TP, _i], _i),
            "basi_tc_conv_set_bn_bwd": ([_P, _TP, _P, _i, _P, _d, _P, _P, _P], _i), "basi_tc_split_kcols": ([_i, _i], _i),
            "basi_avgpool_multi_scratch_floats": ([_TP, _i, _P], C.c_int64),
            "basi_skinny_fwd_workspace_floats": ([_i, _i, _i], C.c_int64)}

_libs = {}
_current = "bf16"


def use(fmt):
    """Selects which build of the library `load()` / `call()` address: "bf16" (default) or "f16".  An Engine binds
    its plan to one build and re-selects it at every entry point, so engines of both formats can coexist."""
    global _current
    if fmt not in LIB_PATHS:
        raise BasiError("unknown library format %r" % (fmt,))
    _current = fmt


def load(fmt=None):
    """dlopen the in-tree library (of the selected 16-bit format); raises BasiError when it has not been built."""
    fmt = fmt or _current
    if fmt in _libs:
        return _libs[fmt]
    path = LIB_PATHS[fmt]
    if not os.path.exists(path):
        raise BasiError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or instance-segment-basi_b200/csrc/build.sh)" % os.path.basename(path))
    lib = C.CDLL(path)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i
    for name, (args, res) in _NOCHECK.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    if lib.basi_half_format() != (1 if fmt == "f16" else 0):
        raise BasiError("%s was built for the other 16-bit format" % path)
    _libs[fmt] = lib
    return lib


def exported_symbols():
    return list(SIGNATURES) + list(_NOCHECK)


LAUNCHES = 0  # number of C-ABI compute calls issued (each enqueues >= 1 of our kernels)


def call(name, *args):
    global LAUNCHES
    lib = load()
    rc = getattr(lib, name)(*args)
    LAUNCHES += 1
    if rc != 0:
        raise BasiError("%s failed (%d): %s" % (name, rc, lib.basi_last_error().decode()))
    return rc


def last_error():
    return load().basi_last_error().decode()


def sm_count():
    return load().basi_sm_count()
