"""ctypes binding of ``libdpi_b200.so`` (the C ABI declared in ``include/dpi_b200.h``).

There is NO fallback: if the shared library is missing the import fails loudly, and every call checks
the return code and raises with ``dpi_last_error_string()``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DPI_B200_LIB") or os.path.join(_HERE, "_C", "libdpi_b200.so")   # (override: A/B experiments)


class DpiError(RuntimeError):
    pass


class ConvGeom(C.Structure):
    """``dpi_conv_geom`` of include/dpi_b200.h."""
    _fields_ = [("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
                ("kd", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32)]


class Parts(C.Structure):
    """``dpi_parts`` of include/dpi_b200.h: a tensor whose channel ranges live in separate dense buffers."""
    _fields_ = [("ptr", C.c_void_p * 4), ("ld", C.c_int64 * 4), ("cbegin", C.c_int32 * 5), ("n", C.c_int32)]

    @staticmethod
    def make(ptrs, lds, widths) -> "Parts":
        t = Parts()
        off = 0
        for i, (p, ld, w) in enumerate(zip(ptrs, lds, widths)):
            t.ptr[i], t.ld[i], t.cbegin[i] = int(p), int(ld), off
            off += int(w)
        t.cbegin[len(widths)] = off
        t.n = len(widths)
        return t


class NextReduce(C.Structure):
    """``dpi_bn_next_reduce`` of include/dpi_b200.h: the BatchNorm-backward reduce of the NEXT unit, fused into the
    apply pass that produces its incoming gradient."""
    _fields_ = [("kind", C.c_int32), ("act", C.c_int32), ("x", Parts), ("mean", C.c_void_p), ("invstd", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("stats_ws", C.c_void_p)]

    @staticmethod
    def make(kind, act, x: Parts, mean, invstd, scale, shift, stats_ws) -> "NextReduce":
        t = NextReduce()
        t.kind, t.act, t.x = int(kind), int(act), x
        t.mean, t.invstd, t.scale, t.shift, t.stats_ws = (int(mean), int(invstd), int(scale) or None, int(shift) or None,
                                                          int(stats_ws))
        return t


class StatsParts(C.Structure):
    """``dpi_stats_parts`` of include/dpi_b200.h: one statistics workspace per channel range."""
    _fields_ = [("ws", C.c_void_p * 4), ("cbegin", C.c_int32 * 5), ("n", C.c_int32)]

    @staticmethod
    def make(ptrs, widths) -> "StatsParts":
        t = StatsParts()
        off = 0
        for i, (p, w) in enumerate(zip(ptrs, widths)):
            t.ws[i], t.cbegin[i] = int(p), off
            off += int(w)
        t.cbegin[len(widths)] = off
        t.n = len(widths)
        return t


class PackJob(C.Structure):
    """``dpi_pack_job`` of include/dpi_b200.h."""
    _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p), ("w_fwd", C.c_void_p), ("w_dgrad", C.c_void_p),
                ("bias_packed", C.c_void_p), ("dw_packed", C.c_void_p), ("dw", C.c_void_p), ("cout_map", C.c_void_p),
                ("cin_map", C.c_void_p), ("Cout_l", C.c_int32), ("Cin_l", C.c_int32), ("Cout_p", C.c_int32),
                ("Cin_p", C.c_int32), ("taps", C.c_int32), ("reserved", C.c_int32), ("w_dgrad_cat", C.c_void_p),
                ("cat_ld", C.c_int32), ("cat_off", C.c_int32), ("cat_taps", C.c_int32), ("cat_tap0", C.c_int32)]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "deep_prior_interpolation_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C deep_prior_interpolation_b200/csrc`). There is no CPU/PyTorch fallback."
            % LIB_PATH)
    return C.CDLL(LIB_PATH)


lib = _load()

_p, _i, _i64, _f, _d, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64
_G = C.POINTER(ConvGeom)
_PT = C.POINTER(Parts)
_NX = C.POINTER(NextReduce)
_SP = C.POINTER(StatsParts)

# name -> (restype, argtypes); must list every symbol of include/dpi_b200.h (tests check this)
SIGNATURES = {
    "dpi_last_error_string": (C.c_char_p, []),
    "dpi_version": (_i, []),
    "dpi_launch_count": (_i64, []),
    "dpi_device_supports_tcgen05": (_i, [_i]),
    "dpi_conv_fwd": (_i, [_p, _i64, _p, _p, _p, _i64, _G, _i, _p]),
    "dpi_conv_fwd_stats": (_i, [_p, _i64, _p, _p, _p, _i64, _G, _i, _p, _p]),
    "dpi_conv_dgrad": (_i, [_p, _i64, _p, _p, _i64, _G, _i, _i, _p]),
    "dpi_conv_dgrad_fused": (_i, [_p, _i64, _p, _p, _i64, _G, _i, _i, _i, _p]),
    "dpi_conv_dgrad_fused_supported": (_i, [_G, _i]),
    "dpi_conv_wgrad_workspace_bytes": (_i64, [_G]),
    "dpi_conv_wgrad": (_i, [_p, _i64, _p, _i64, _p, _G, _p, _i64, _i, _p]),
    "dpi_pack_conv_weights": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p]),
    "dpi_unpack_conv_wgrad": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "dpi_pack_conv_weights_batched": (_i, [_p, _i, _i, _p]),
    "dpi_channel_stats_parts": (_i, [_PT, _i64, _i, _p, _p]),
    "dpi_add_affine_act_parts": (_i, [_p, _i64, _PT, _p, _p, _p, _i, _p, _i64, _i64, _i, _p, _p]),
    "dpi_bn_bwd_reduce_parts": (_i, [_p, _i64, _p, _i64, _i, _PT, _p, _p, _i64, _i, _p, _p]),
    "dpi_bn_bwd_apply_parts": (_i, [_p, _i64, _p, _i64, _i, _PT, _p, _p, _p, _p, _p, _PT, _i, _p, _i64, _i64, _i, _p]),
    "dpi_unpack_conv_wgrad_batched": (_i, [_p, _i, _p]),
    "dpi_bias_grad": (_i, [_p, _i64, _i64, _i, _p, _p, _p, _i64, _p]),
    "dpi_stats_workspace_bytes": (_i64, [_i]),
    "dpi_channel_stats": (_i, [_p, _i64, _i64, _i, _p, _p]),
    "dpi_bn_finalize": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p]),
    "dpi_bn_finalize_parts": (_i, [_SP, _i64, _i, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p]),
    "dpi_affine_act": (_i, [_p, _i64, _p, _p, _p, _i, _p, _i64, _i64, _i, _p, _p]),
    "dpi_add_affine_act": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _i, _p, _i64, _i64, _i, _p, _p]),
    "dpi_act_bwd": (_i, [_p, _i64, _p, _i64, _i, _p, _i64, _i64, _i, _i, _p]),
    "dpi_bn_bwd_reduce": (_i, [_p, _i64, _p, _i64, _i, _p, _i64, _p, _p, _p, _p, _i64, _i, _p, _p]),
    "dpi_bn_bwd_finalize": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "dpi_bn_bwd_apply": (_i, [_p, _i64, _p, _i64, _i, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _i, _p]),
    "dpi_bn_bwd_apply_next": (_i, [_p, _i64, _p, _i64, _i, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _i, _NX, _p]),
    "dpi_bn_bwd_apply_parts_next": (_i, [_p, _i64, _p, _i64, _i, _PT, _p, _p, _p, _p, _p, _PT, _i, _p, _i64, _i64, _i, _NX,
                                         _p]),
    "dpi_upsample2x_fwd": (_i, [_p, _i64, _i, _i, _i, _p, _i64, _i, _i, _i, _i, _i, _i, _p]),
    "dpi_upsample2x_bwd": (_i, [_p, _i64, _i, _i, _i, _p, _i64, _i, _i, _i, _i, _i, _i, _i, _p]),
    "dpi_copy_slice": (_i, [_p, _i64, _p, _i64, _i64, _i, _i, _p]),
    "dpi_gate_mul_fwd": (_i, [_p, _i64, _p, _i64, _p, _i64, _i64, _i, _i, _p]),
    "dpi_gate_mul_bwd": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i, _i, _p]),
    "dpi_nchw_to_cl": (_i, [_p, _i, _i64, _p, _p, _i64, _i, _p]),
    "dpi_cl_to_nchw": (_i, [_p, _i64, _i, _p, _p, _i, _i64, _p]),
    "dpi_noise_axpy": (_i, [_p, _p, _p, _i64, _f, _u64, _u64, _i, _p]),
    "dpi_noise_axpy_dev": (_i, [_p, _p, _i64, _f, _u64, _p, _i, _p]),
    "dpi_add_data_dev": (_i, [_p, _i64, _i, _i64, _p, _i64, _i, _p, _i, _p, _i, _p]),
    "dpi_fir_axis": (_i, [_p, _p, _i64, _i64, _i64, _p, _i, _p]),
    "dpi_iteration_end": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p, _i64, _p]),
    "dpi_fill_normal": (_i, [_p, _i64, _f, _f, _u64, _u64, _p]),
    "dpi_loss_workspace_bytes": (_i64, []),
    "dpi_masked_loss": (_i, [_p, _p, _p, _i64, _i64, _i, _p, _p, _i64, _p, _p]),
    "dpi_adam_step": (_i, [_p, _p, _p, _p, _i64, _d, _d, _d, _d, _d, _i64, _p]),
    "dpi_axpby": (_i, [_p, _p, _f, _f, _i64, _p]),
    "dpi_adam_step_dev": (_i, [_p, _p, _p, _p, _i64, _p, _d, _d, _d, _d, _p]),
    "dpi_patch_extract_f64": (_i, [_p, _p, _p, _p, _d, _p, _p]),
    "dpi_patch_reassemble_f32": (_i, [_p, _p, _p, _p, _f, _p, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

_INT_RETURNING = {n for n, (r, _) in SIGNATURES.items() if r is _i and n not in ("dpi_version", "dpi_device_supports_tcgen05")}


def last_error() -> str:
    return lib.dpi_last_error_string().decode("utf-8", "replace")


def check(rc: int, name: str = "dpi") -> None:
    if rc != 0:
        raise DpiError("%s failed (rc=%d): %s" % (name, rc, last_error()))


def call(name: str, *args):
    """Call an int-returning entry point and raise DpiError on a non-zero return code."""
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise DpiError("%s failed (rc=%d): %s" % (name, rc, last_error()))


# constants of include/dpi_b200.h
ACT_CODES = {None: 0, "none": 0, "LeakyReLU": 1, "ReLU": 2, "ELU": 3, "Tanh": 4, "Sigmoid": 5}
PREC_FP32, PREC_TF32 = 0, 1
LOSS_CODES = {"mae": 0, "mse": 1}
UP_NEAREST, UP_LINEAR = 0, 1
ROUND_TF32 = 0x100
