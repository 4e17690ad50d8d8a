"""ctypes binding of libunmicst_b200.so (include/unmicst_b200.h).

There is deliberately no fallback: if the CUDA library is missing or does not
load, importing the engine raises.  Set UNMICST_B200_AUTOBUILD=1 to compile it
on first use (needs nvcc)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UNMICST_B200_LIB") or os.path.join(HERE, "csrc", "libunmicst_b200.so")     # (override: A/B builds of tools/)

UMX_ABI_VERSION = 1
UMX_OK, UMX_EINVAL, UMX_ENOTENSOR, UMX_ECUDA, UMX_ENOMEM, UMX_ENODEVICE = 0, -1, -2, -3, -4, -5
UMX_GRAPH_LEGACY, UMX_GRAPH_V2 = 0, 1
UMX_U8, UMX_U16, UMX_F32, UMX_F64 = 0, 1, 2, 3
UMX_PREC_DEFAULT, UMX_PREC_FP32, UMX_PREC_SPLIT3, UMX_PREC_SINGLE, UMX_PREC_MIXED = 0, 1, 2, 3, 4
UMX_F_NO_SYNC = 1
UMX_F_CLI_QUANT = 2
UMX_F_PREMAP_PER_PLANE = 4
UMX_F_CONTINUE = 8
UMX_F_STITCH_REPLACE = 16
UMX_F_FP16_QUANT = 32

PRECISIONS = {"default": UMX_PREC_DEFAULT, "fp32": UMX_PREC_FP32, "split3": UMX_PREC_SPLIT3, "single": UMX_PREC_SINGLE,
              "mixed": UMX_PREC_MIXED}


class umx_model_desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "graph", "im_size", "n_channels", "n_classes", "n_out0", "n_layers", "feat_maps_fact",
        "down_samp_fact", "ks", "n_extra_convs", "precision", "max_batch_tiles")] + [("reserved", C.c_int32 * 3)]


class umx_tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class umx_premap(C.Structure):
    _fields_ = [("in_scale", C.c_double), ("rescale", C.c_int32), ("pad_", C.c_int32),
                ("imin", C.c_double), ("imax", C.c_double), ("omin", C.c_double), ("omax", C.c_double)]


class umx_opts(C.Structure):
    _fields_ = [("tile_row0", C.c_int32), ("tile_row1", C.c_int32), ("precision", C.c_int32), ("flags", C.c_int32),
                ("premap", C.POINTER(umx_premap)), ("out_plane_stride", C.c_int64), ("out_row_base", C.c_int32),
                ("infer_h", C.c_int32), ("infer_w", C.c_int32), ("reserved", C.c_int32 * 3)]


class umx_image(C.Structure):
    _fields_ = [("img", C.c_void_p), ("dtype", C.c_int32), ("n_planes", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("plane_stride", C.c_int64), ("premap", C.POINTER(umx_premap)), ("out_u8", C.c_void_p), ("out_f32", C.c_void_p)]


class umx_prof_entry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("ms_total", C.c_double),
                ("flops", C.c_double), ("bytes", C.c_double)]


# every symbol include/unmicst_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "umx_device_count": (C.c_int, []),
    "umx_device_free_mem": (C.c_int, [C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "umx_create": (C.c_int, [C.POINTER(umx_model_desc), C.POINTER(umx_tensor), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "umx_create_ex": (C.c_int, [C.POINTER(umx_model_desc), C.POINTER(umx_tensor), C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                C.POINTER(C.c_void_p)]),
    "umx_set_op_terms": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "umx_destroy": (None, [C.c_void_p]),
    "umx_forward_tiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
    "umx_infer_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                  C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.POINTER(umx_opts)]),
    "umx_infer_images": (C.c_int, [C.c_void_p, C.POINTER(umx_image), C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "umx_band_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "umx_band_out_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "umx_resample_minmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "umx_set_stream": (C.c_int, [C.c_void_p, C.c_uint64]),
    "umx_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "umx_profile_read": (C.c_int, [C.c_void_p, C.POINTER(umx_prof_entry), C.c_int32, C.c_int32]),
    "umx_launch_count": (C.c_int64, [C.c_void_p]),
    "umx_op_info": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32]),
    "umx_debug_buffer": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_int64]),
    "umx_host_alloc": (C.c_void_p, [C.c_int64]),
    "umx_host_free": (None, [C.c_void_p]),
    "umx_describe_plan": (C.c_int64, [C.POINTER(umx_model_desc), C.POINTER(umx_tensor), C.c_int32, C.c_char_p, C.c_int64]),
    "umx_tiff_lzw_decode": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "umx_last_error": (C.c_char_p, []),
    "umx_version": (C.c_char_p, []),
}

_lib = None


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"unmicst_b200 error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load the CUDA library once; raise (never fall back) if it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and os.environ.get("UNMICST_B200_AUTOBUILD") == "1":
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built; run `python -m unmicst_b200.build` (needs nvcc). "
                          "unmicst_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(code: int) -> None:
    if code != UMX_OK:
        raise EngineError(code, lib().umx_last_error().decode("utf-8", "replace"))
