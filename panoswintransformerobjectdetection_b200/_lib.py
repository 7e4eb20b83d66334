"""ctypes binding of libpanoswin_b200.so (include/panoswin_b200.h).  The library is the product:
if it is missing or cannot be loaded this module raises — there is no Python / CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

PSW_F32, PSW_BF16 = 0, 1
PSW_EPI_GELU = 1

_vp, _fp, _i, _i64, _f = C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float

ABI_VERSION = 6            # must equal PSW_ABI_VERSION of include/panoswin_b200.h (bumped with every prototype change)

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
    "psw_window_attn_fwd": [_vp, _vp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp],
    "psw_window_grid": [_i, _i, _i, _i, _vp, _vp],
    "psw_window_source_map": [_i, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "psw_window_attn_full_supported": [_i, _i],
    "psw_window_bias_full_bytes": [_i, _i, _i, _i, _i],
    "psw_window_bias_full": [_fp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "psw_window_attn_full_fwd": [_vp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp],
    "psw_patch_merge_ln_fwd": [_vp, _vp, _fp, _fp, _i, _i, _i, _i, _f, _i, _i, _vp],
    "psw_layernorm_nchw_fwd": [_vp, _vp, _fp, _fp, _i, _i64, _i, _f, _i, _vp],
    "psw_stem_conv3x3_relu_fwd": [_fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _vp],
    "psw_stem_conv3x3_c32_relu_fwd": [_vp, _vp, _fp, _vp, _i, _i, _i, _i, _vp],
    "psw_patch_conv_fwd": [_vp, _vp, _fp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "psw_conv3x3_nhwc_fwd": [_vp, _vp, _fp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "psw_cast": [_vp, _vp, _i64, _i, _i, _vp],
    "psw_conv2d_f32_fwd": [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "psw_layernorm_bwd": [_vp, _vp, _fp, _vp, _fp, _fp, _fp, _i64, _i, _f, _i, _i, _vp],
    "psw_patch_merge_ln_bwd": [_vp, _vp, _fp, _vp, _fp, _fp, _fp, _i, _i, _i, _i, _f, _i, _i, _vp],
    "psw_linear_bwd_workspace_bytes": [_i64, _i, _i, _i],
    "psw_linear_bwd": [_vp, _vp, _vp, _vp, _fp, _fp, _i64, _i, _i, _i, _i, _vp, _i64, _vp],
    "psw_bn_stats_fwd": [_vp, _fp, _fp, _i64, _i, _vp],
    "psw_bn_apply_relu_fwd": [_vp, _vp, _fp, _fp, _i64, _i, _vp],
    "psw_bn_relu_bwd": [_vp, _vp, _vp, _fp, _fp, _fp, _vp, _fp, _fp, _i64, _i, _vp],
    "psw_gelu_fwd": [_vp, _vp, _i64, _i, _vp],
    "psw_gelu_bwd": [_vp, _vp, _vp, _i64, _i, _vp],
    "psw_transpose": [_vp, _vp, _i64, _i64, _i, _vp],
    "psw_window_attn_bwd": [_vp, _vp, _fp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp],
}
# include/panoswin_b200_debug.h: present only in the -DPSW_DIAGNOSTICS build (load(diagnostics=True))
DIAG_SIGNATURES = {
    "psw_diag_window_attn_full": [_vp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _i, _vp],
    "psw_diag_linear_mode": [_i],
    "psw_diag_linear_cycles": [_vp],
    "psw_diag_mlp_mode": [_i],
}

_libs = {}
# tools/microbench.py sets PSW_DIAGNOSTICS=1 before importing the package: every call then goes to the diagnostics build
_DEFAULT_DIAG = os.environ.get("PSW_DIAGNOSTICS") == "1"


class PanoSwinB200Error(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True, diagnostics: bool = None) -> C.CDLL:
    """Load and type the library.  When nvcc is available the (cheap) source fingerprint is checked first and a stale
    or missing library is rebuilt (file-locked: torchrun ranks do not race); without nvcc a stale stamp raises.  The
    ABI version of the loaded binary must equal ABI_VERSION (a stale .so called through new argtypes would corrupt
    memory silently).  `diagnostics=True` loads the -DPSW_DIAGNOSTICS build used by the profiling tools."""
    if diagnostics is None:
        diagnostics = _DEFAULT_DIAG
    if diagnostics in _libs:
        return _libs[diagnostics]
    path = _build.DIAG_LIB_PATH if diagnostics else library_path()
    if not _build.is_current(diagnostics):
        have_nvcc = _build._nvcc(required=False) is not None
        if have_nvcc and build_if_missing:
            _build.build(diagnostics=diagnostics)
        elif not os.path.isfile(path):
            raise PanoSwinB200Error(f"{path} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        elif not have_nvcc:
            raise PanoSwinB200Error(f"{path} does not match the sources (stale build) and nvcc is not available to rebuild it")
    lib = C.CDLL(path)
    lib.psw_abi_version.argtypes = []
    lib.psw_abi_version.restype = C.c_int
    if lib.psw_abi_version() != ABI_VERSION:
        raise PanoSwinB200Error(f"{path}: ABI version {lib.psw_abi_version()} != {ABI_VERSION} expected by the Python binding")
    sigs = dict(SIGNATURES)
    if diagnostics:
        sigs.update(DIAG_SIGNATURES)
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)            # AttributeError here = header / library mismatch
        fn.argtypes = argtypes
        fn.restype = C.c_int64 if name in ("psw_window_bias_full_bytes", "psw_linear_bwd_workspace_bytes") else C.c_int
    lib.psw_last_error_string.argtypes = []
    lib.psw_last_error_string.restype = C.c_char_p
    _libs[diagnostics] = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().psw_last_error_string().decode("utf-8", "replace")
        kind = "argument/unsupported" if rc < 0 else "CUDA"
        raise PanoSwinB200Error(f"{what} failed ({kind} error {rc}): {msg}")
