"""ctypes binding of ``libmermaid_b200.so`` (the C ABI declared in ``include/mermaid_b200.h``).

There is no CPU fallback: if the shared library is missing and cannot be built, or if a
compute entry point is called without a CUDA device, this module raises.
"""

from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libmermaid_b200.so"
HEADER = PKG.parent / "include" / "mermaid_b200.h"

MC_OK = 0
MC_ERR_BAD_ARG = 1
MC_ERR_POINT_BOUNDS = 2
MC_ERR_DATA_LIMIT = 3
MC_ERR_CUDA = 4
MC_ERR_UNSUPPORTED = 5
MC_ERR_NOMEM = 6

MODE_FP32 = 0
MODE_BF16 = 1
MODES = {"fp32": MODE_FP32, "bf16": MODE_BF16}
FEATURE_DIM = 1280
CROP_SIZE = 224


class RowColumnInvalidError(ValueError):
    """Mirror of ``spacer.exceptions.RowColumnInvalidError`` (raised by check_extract_inputs)."""


class DataLimitError(ValueError):
    """Mirror of ``spacer.exceptions.DataLimitError`` (raised by check_extract_inputs)."""


class McImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("height", C.c_int32), ("width", C.c_int32), ("row_pitch", C.c_int64)]


class McPoint(C.Structure):
    _fields_ = [("image", C.c_int32), ("row", C.c_int32), ("col", C.c_int32)]


GRAD_SYNC_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)

_vp, _i32, _i64, _u32, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every symbol include/mermaid_b200.h declares
SIGNATURES = {
    "mc_abi_version": (C.c_int, []),
    "mc_last_error": (C.c_char_p, []),
    "mc_synth_image": (C.c_int, [_vp, _i32, _i32, _i64, _u32, _u32, _vp]),
    "mc_check_extract_inputs": (C.c_int, [_i32, _i32, _vp, _i64, _i64, _i64]),
    "mc_crop_patches": (C.c_int, [_vp, _i32, _vp, _i64, _vp, _vp]),
    "mc_crop_resize_patches": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp, _vp]),
    "mc_normalize_patches": (C.c_int, [_vp, _i64, _vp, _vp]),
    "mc_backbone_param_count": (_i64, []),
    "mc_extractor_create": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _pp]),
    "mc_extractor_destroy": (C.c_int, [_vp]),
    "mc_extractor_mode": (C.c_int, [_vp]),
    "mc_extractor_launches": (_i64, [_vp]),
    "mc_extract_points": (C.c_int, [_vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "mc_extract_patches": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "mc_extract_image_host": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _vp, _i64, _vp, _vp]),
    "mc_extract_images_host": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp]),
    "mc_extractor_pipe_stats": (C.c_int, [_vp, _vp, _vp, _vp]),
    "mc_upload_window": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "mc_plan_uploads": (C.c_int, [_i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "mc_jpeg_create": (C.c_int, [_i32, _pp]),
    "mc_jpeg_destroy": (C.c_int, [_vp]),
    "mc_jpeg_info": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "mc_jpeg_decode": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _vp]),
    "mc_jpeg_decode_exact": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _vp]),
    "mc_jpeg_coefficients_host": (C.c_int, [_vp, _i64, _vp, _i64, _vp]),
    "mc_extractor_set_tap": (C.c_int, [_vp, _i32, _vp, _i64]),
    "mc_extractor_profile": (C.c_int, [_vp, _i32]),
    "mc_extractor_profile_read": (C.c_int, [_vp, _vp, _vp, _i32]),
    "mc_head_create": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _i32, _pp]),
    "mc_head_destroy": (C.c_int, [_vp]),
    "mc_head_set_exact": (C.c_int, [_vp, _i32]),
    "mc_head_scores": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "mc_head_scores_host": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "mc_head_launches": (_i64, [_vp]),
    "mc_head_evaluate": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "mc_platt_fit": (C.c_int, [_vp, _vp, _i64, _i32, _i32, C.c_double, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mc_mlp_create": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _i32, _pp]),
    "mc_mlp_destroy": (C.c_int, [_vp]),
    "mc_dp_unique_id": (C.c_int, [_vp]),
    "mc_dp_create": (C.c_int, [_vp, _i32, _i32, _i32, _pp]),
    "mc_dp_destroy": (C.c_int, [_vp]),
    "mc_dp_all_reduce_sum": (C.c_int, [_vp, _vp, _i64, _vp]),
    "mc_mlp_partial_fit": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mc_mlp_get_params": (C.c_int, [_vp, _vp, _vp]),
    "mc_mlp_get_adam": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mc_mlp_set_adam": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64]),
    "mc_mlp_steps": (_i64, [_vp]),
    "mc_mlp_launches": (_i64, [_vp]),
    "mc_mlp_graph_steps": (_i64, [_vp]),
    "mc_mlp_grad_size": (_i64, [_vp]),
}

_lib = None


def header_symbols() -> list[str]:
    """Function names declared in include/mermaid_b200.h."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", text)) - {"mc_grad_sync_fn"})


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load the shared library (building it with nvcc if it is not there yet)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    if not LIB_PATH.exists() or _build.is_stale():
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing or stale; run `python -m mermaid_classifier_b200.build`")
        _build.build()
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mc_abi_version() != 1:
        raise RuntimeError("libmermaid_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def last_error() -> str:
    return load().mc_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    """Map a C status to the exception type the reference raises at the same place."""
    if status == MC_OK:
        return
    msg = last_error()
    if status == MC_ERR_BAD_ARG:
        raise ValueError(msg)
    if status == MC_ERR_POINT_BOUNDS:
        raise RowColumnInvalidError(msg)
    if status == MC_ERR_DATA_LIMIT:
        raise DataLimitError(msg)
    if status == MC_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(f"libmermaid_b200 error {status}: {msg}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "mermaid_classifier_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback"
        )
    return torch


def stream_ptr(torch_stream=None) -> int:
    import torch

    s = torch_stream if torch_stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)
