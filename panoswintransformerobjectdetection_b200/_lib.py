"""ctypes binding of libpanoswin_b200.so (include/panoswin_b200.h).  The library is the product:
if it is missing or cannot be loaded this module raises — there is no Python / CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

PSW_F32, PSW_BF16 = 0, 1
PSW_EPI_GELU = 1

_vp, _fp, _i, _i64, _f = C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes, exactly the prototypes of include/panoswin_b200.h
SIGNATURES = {
    "psw_abi_version": [],
    "psw_check_device": [_i],
    "psw_layernorm_fwd": [_vp, _vp, _fp, _fp, _fp, _i64, _i, _i64, _f, _i, _i, _vp],
    "psw_layernorm2_fwd": [_vp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _i64, _i, _i64, _f, _f, _i, _vp],
    "psw_linear_fwd": [_vp, _vp, _fp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "psw_linear_ln_fwd": [_vp, _vp, _fp, _vp, _vp, _fp, _fp, _f, _vp, _i64, _i, _i, _vp],
    "psw_linear_ln_nchw_fwd": [_vp, _vp, _fp, _vp, _vp, _fp, _fp, _f, _fp, _i64, _i64, _i, _i, _vp],
    "psw_mlp_fused_fwd": [_vp, _vp, _fp, _vp, _fp, _vp, _i64, _i, _i, _vp],
    "psw_window_attn_fwd": [_vp, _vp, _fp, _fp, _vp, _fp, _fp, _vp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp],
    "psw_window_bias_tables": [_fp, _fp, _vp, _i, _i, _vp],
    "psw_window_grid": [_i, _i, _i, _i, _vp, _vp],
    "psw_window_hav_table": [_fp, _vp, _i, _i, _i, _i, _vp],
    "psw_window_bias_full_bytes": [_i, _i, _i, _i, _i],
    "psw_window_bias_full": [_fp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "psw_window_attn_full_fwd": [_vp, _vp, _vp, _fp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp],
    "psw_patch_merge_ln_fwd": [_vp, _vp, _fp, _fp, _i, _i, _i, _i, _f, _i, _i, _vp],
    "psw_layernorm_nchw_fwd": [_vp, _vp, _fp, _fp, _i, _i64, _i, _f, _i, _vp],
    "psw_stem_conv3x3_relu_fwd": [_fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _vp],
    "psw_stem_conv3x3_c32_relu_fwd": [_vp, _vp, _fp, _vp, _i, _i, _i, _i, _vp],
    "psw_patch_conv_fwd": [_vp, _vp, _fp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "psw_patch_conv_ln_fwd": [_vp, _vp, _fp, _fp, _fp, _f, _fp, _i64, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "psw_cast": [_vp, _vp, _i64, _i, _i, _vp],
    "psw_window_attn_fwd_profile": [_vp, _vp, _fp, _fp, _vp, _fp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _vp],
    "psw_debug_linear_mode": [_i],
    "psw_debug_mlp_mode": [_i],
    "psw_debug_linear_cycles": [_vp],
    "psw_debug_source_map": [_i, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "psw_window_attn_fwd_simt_bf16": [_vp, _vp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp],
}

_lib = None


class PanoSwinB200Error(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first when the .so is absent and nvcc is available) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.isfile(path):
        if not build_if_missing:
            raise PanoSwinB200Error(f"{path} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        _build.build()
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError here = header / library mismatch
        fn.argtypes = argtypes
        fn.restype = C.c_int64 if name == "psw_window_bias_full_bytes" else C.c_int
    lib.psw_last_error_string.argtypes = []
    lib.psw_last_error_string.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().psw_last_error_string().decode("utf-8", "replace")
        kind = "argument/unsupported" if rc < 0 else "CUDA"
        raise PanoSwinB200Error(f"{what} failed ({kind} error {rc}): {msg}")
