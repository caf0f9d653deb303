"""ctypes binding of libconan_b200.so (the C ABI declared in include/conan_b200.h).

The library is built in-tree by `python -m conan_b200.build` (or __graft_entry__.build()).
There is no fallback: if the shared object is missing or cannot be loaded, importing the
product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libconan_b200.so")
ABI_VERSION = 1

DTYPE_F32, DTYPE_F16, DTYPE_I32 = 0, 1, 2


class ConanConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("max_slots", C.c_int32), ("max_ref_frames", C.c_int32),
        ("emformer_layers", C.c_int32), ("emformer_dim", C.c_int32), ("emformer_heads", C.c_int32),
        ("emformer_ffn", C.c_int32), ("segment", C.c_int32), ("right_context", C.c_int32),
        ("left_context", C.c_int32), ("emformer_output_dim", C.c_int32),
        ("hidden_size", C.c_int32), ("content_kernel", C.c_int32), ("dec_blocks", C.c_int32),
        ("dec_kernel", C.c_int32), ("dec_post_kernel", C.c_int32), ("predictor_kernel", C.c_int32),
        ("n_vq", C.c_int32), ("silent_token", C.c_int32), ("n_mels", C.c_int32),
        ("voc_initial_channel", C.c_int32), ("voc_n_ups", C.c_int32), ("voc_rates", C.c_int32 * 8),
        ("voc_up_kernels", C.c_int32 * 8), ("voc_n_res", C.c_int32), ("voc_res_kernels", C.c_int32 * 8),
        ("voc_res_dilations", C.c_int32 * 8), ("voc_n_dil", C.c_int32),
        ("voc_precision", C.c_int32), ("voc_use_tensor_cores", C.c_int32), ("voc_group", C.c_int32),
        ("voc_residual_from_ctx", C.c_int32), ("lin_use_tensor_cores", C.c_int32), ("voc_fuse_resblocks", C.c_int32), ("lin_fuse_ffn", C.c_int32),
        ("ses_use_tensor_cores", C.c_int32), ("emformer_memory_size", C.c_int32), ("step_graphs", C.c_int32), ("lin_fuse_blocks", C.c_int32),
    ]


class ConvParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_slot_stride", C.c_int64), ("x_row_stride", C.c_int32), ("x_rows", C.c_int32),
        ("x_is_half", C.c_int32), ("row0", C.c_int32), ("L", C.c_int32),
        ("cin", C.c_int32), ("k", C.c_int32), ("dil", C.c_int32), ("cout", C.c_int32),
        ("w", C.c_void_p), ("bias", C.c_void_p), ("n_streams", C.c_int32), ("slot_ids", C.c_void_p),
        ("n_slots", C.c_int32), ("scale", C.c_float), ("act", C.c_int32), ("slope", C.c_float),
        ("res", C.c_void_p), ("res_slot_stride", C.c_int64), ("res_row_stride", C.c_int32),
        ("rowmask", C.c_void_p), ("mask_slot_stride", C.c_int32), ("out_scale", C.c_float),
        ("y", C.c_void_p), ("y_slot_stride", C.c_int64), ("y_row_stride", C.c_int32), ("y_row0", C.c_int32),
        ("accumulate", C.c_int32),
        ("y2", C.c_void_p), ("y2_slot_stride", C.c_int64), ("y2_row_stride", C.c_int32), ("y2_row0", C.c_int32),
        ("y2_is_half", C.c_int32), ("act2", C.c_int32), ("slope2", C.c_float),
        ("x_split", C.c_int32), ("x_lo_slot_off", C.c_int64), ("acc_scale", C.c_float),
        ("y2_split", C.c_int32), ("y2_lo_off", C.c_int64),
        ("res_is_half", C.c_int32), ("res_inv_slope", C.c_float),
        ("res2", C.c_void_p), ("res2_slot_stride", C.c_int64), ("res2_row_stride", C.c_int32), ("res2_is_half", C.c_int32),
        ("y_is_half", C.c_int32),
    ]


# every symbol include/conan_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "conan_last_error": (C.c_char_p, []),
    "conan_abi_version": (C.c_int, []),
    "conan_sizeof_config": (C.c_size_t, []),
    "conan_sizeof_conv_params": (C.c_size_t, []),
    "conan_engine_create": (C.c_int, [C.POINTER(ConanConfig), C.POINTER(_P)]),
    "conan_engine_destroy": (None, [_P]),
    "conan_engine_num_weights": (C.c_int, [_P]),
    "conan_engine_weight_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "conan_engine_bind_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_size_t, C.c_int]),
    "conan_engine_finalize": (C.c_int, [_P]),
    "conan_engine_state_bytes": (C.c_size_t, [_P]),
    "conan_slots_reset": (C.c_int, [_P, C.c_int, _P, C.c_int, _P]),
    "conan_session_open": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P]),
    "conan_emformer_step": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "conan_emformer_forward": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, _P, _P, _P]),
    "conan_decoder_step": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "conan_vocoder_step": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "conan_step": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "conan_step_host": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "conan_step_host_submit": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_int)]),
    "conan_step_host_wait": (C.c_int, [_P, C.c_int]),
    "conan_engine_launch_count": (C.c_uint64, [_P]),
    "conan_engine_graph_replays": (C.c_uint64, [_P]),
    "conan_engine_set_profiling": (C.c_int, [_P, C.c_int]),
    "conan_engine_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "conan_debug_read": (C.c_int, [_P, C.c_char_p, C.c_int, _P, C.c_size_t, C.POINTER(C.c_size_t), _P]),
    "conan_conv_gemm": (C.c_int, [C.POINTER(ConvParams), C.c_int, _P]),
    "conan_logmel": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_float,
                               C.c_float, C.c_float, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the shared library and sets the prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m conan_b200.build` "
            "(conan_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.conan_abi_version() != ABI_VERSION:
        raise RuntimeError("libconan_b200.so ABI version mismatch; rebuild")
    if lib.conan_sizeof_config() != C.sizeof(ConanConfig) or lib.conan_sizeof_conv_params() != C.sizeof(ConvParams):
        raise RuntimeError("ctypes struct layout does not match include/conan_b200.h; fix conan_b200/_lib.py")
    _lib = lib
    return lib


def last_error() -> str:
    return load().conan_last_error().decode()


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"conan_b200 {what} failed: {last_error()}")
