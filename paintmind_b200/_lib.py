"""ctypes binding of libpaintmind_b200.so (the C-ABI in include/paintmind_b200.h).

The product path has NO fallback: if the shared library is missing, or a call returns non-zero,
a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# PM_B200_LIB: another build of the same library (same-box A/B of two builds, e.g. scripts/abwd_opaque_ab.sh); default: the in-tree one
_LIB_PATH = Path(os.environ.get("PM_B200_LIB") or (Path(__file__).resolve().parent / "lib" / "libpaintmind_b200.so"))
_lib = None

PM_OUT_BF16, PM_OUT_F32, PM_OUT_UNPATCH, PM_OUT_UNPATCH_U8 = 0, 1, 2, 3


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("out", C.c_void_p),
        ("bias", C.c_void_p), ("colsum", C.c_void_p), ("stats", C.c_void_p),
        ("pos", C.c_void_p), ("res", C.c_void_p),
        ("lda", C.c_int64), ("ldw", C.c_int64), ("ld_out", C.c_int64),
        ("ld_pos", C.c_int64), ("ld_res", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("pos_rows", C.c_int32), ("out_mode", C.c_int32), ("swiglu", C.c_int32), ("bn", C.c_int32),
        ("patch", C.c_int32), ("channels", C.c_int32), ("grid", C.c_int32), ("max_ctas", C.c_int32),
        ("stats_raw", C.c_int32), ("ln_eps", C.c_float), ("stats_out", C.c_void_p),
        ("cta_group", C.c_int32), ("debug", C.c_void_p), ("res_mod", C.c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64), ("ldo", C.c_int64),
        ("bsq", C.c_int64), ("bsk", C.c_int64), ("bsv", C.c_int64), ("bso", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Nq", C.c_int32), ("Nk", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float),
        ("lse", C.c_void_p), ("lse_ld", C.c_int64), ("o32", C.c_void_p), ("ldo32", C.c_int64),
        ("q_prescaled", C.c_int32), ("reserved", C.c_int32),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p), ("d_o", C.c_void_p),
        ("lse", C.c_void_p), ("delta", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64), ("ldo", C.c_int64), ("lddo", C.c_int64),
        ("lddq", C.c_int64), ("lddk", C.c_int64), ("lddv", C.c_int64),
        ("bsq", C.c_int64), ("bsk", C.c_int64), ("bsv", C.c_int64), ("bso", C.c_int64), ("bsdo", C.c_int64),
        ("bsdq", C.c_int64), ("bsdk", C.c_int64), ("bsdv", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Nq", C.c_int32), ("Nk", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float), ("o_is_f32", C.c_int32), ("lse_ld", C.c_int64), ("debug", C.c_void_p),
    ]


class VqArgs(C.Structure):
    _fields_ = [
        ("z", C.c_void_p), ("en", C.c_void_p), ("packed", C.c_void_p),
        ("cand_val", C.c_void_p), ("cand_idx", C.c_void_p),
        ("idx", C.c_void_p), ("zq", C.c_void_p), ("zq_split", C.c_void_p),
        ("sse", C.c_void_p), ("hist", C.c_void_p),
        ("ldz", C.c_int64),
        ("M", C.c_int32), ("n_e", C.c_int32), ("e_dim", C.c_int32), ("splits", C.c_int32),
    ]


class MaskgitSampleArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("noise", C.c_void_p), ("ids", C.c_void_p), ("pred_ids", C.c_void_p),
        ("scores", C.c_void_p),
        ("ld", C.c_int64), ("ld_noise", C.c_int64), ("mask_id", C.c_int64),
        ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("M", C.c_int32), ("V", C.c_int32), ("topk", C.c_int32), ("temperature", C.c_float),
        ("step_tab", C.c_void_p), ("step_idx", C.c_void_p),
    ]


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} not found: build it with `python -m paintmind_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(str(_LIB_PATH))
    lib.pm_version.restype = C.c_int
    lib.pm_device_check.restype = C.c_int
    lib.pm_error_string.restype = C.c_char_p
    lib.pm_error_string.argtypes = [C.c_int]
    for name in WORKSPACE_QUERIES:
        if not hasattr(lib, name):
            raise RuntimeError(f"libpaintmind_b200.so does not export {name}")
        fn = getattr(lib, name)
        fn.restype = C.c_int64
        fn.argtypes = [C.c_int32] * WORKSPACE_QUERIES[name]
    for name, argtypes in EXPORTS.items():
        if not hasattr(lib, name):
            raise RuntimeError(f"libpaintmind_b200.so does not export {name}")
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


# every compute entry point declared in include/paintmind_b200.h, with its C signature
_p, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
EXPORTS = {
    "pm_gemm_bf16": [C.POINTER(GemmArgs), _p],
    "pm_attn_fwd": [C.POINTER(AttnArgs), _p],
    "pm_vq_codebook_prep": [_p, _i32, _i32, _p, _p, _p],
    "pm_vq_fwd": [C.POINTER(VqArgs), _p],
    "pm_vq_gather": [_p, _i32, _i32, _i32, _p, _i32, _p, _p, _p],
    "pm_split_rows32": [_p, _i64, _i32, _p, _p],
    "pm_patchify8": [_p, _p, _i32, _i32, _i32, _i32, _p],
    "pm_patchify8_u8": [_p, _p, _i32, _i32, _i32, _p],
    "pm_layernorm": [_p, _i64, _i32, _i32, _f, _p, _p, _p, _i64, _p, _p],
    "pm_cast_f32_bf16": [_p, _p, _i64, _p],
    "pm_maskgit_sample": [C.POINTER(MaskgitSampleArgs), _p],
    "pm_maskgit_remask": [_p, _p, _i32, _i32, _i32, _i64, _p],
    "pm_maskgit_remask_step": [_p, _p, _i32, _i32, _i64, _p, _p, _p, _p],
    "pm_maskgit_random_mask": [_p, _i64, _p, C.c_uint64, C.c_uint64, _p, _i32, _i32, _i32, _p, _p, _p],
    "pm_ce_label_smooth": [_p, _i64, _i32, _i32, _p, _p, _f, _p, _p, _p, _p],
    # generator backward path
    "pm_attn_bwd": [C.POINTER(AttnBwdArgs), _p],
    "pm_wgrad_bf16": [_p, _i64, _p, _i64, _i32, _i32, _i32, _p, _p, _i64, _i32, _p],
    "pm_colsum_bf16": [_p, _i64, _i32, _i32, _p, _p, _i32, _p],
    "pm_layernorm_bwd": [_p, _i64, _p, _i64, _p, _p, _i64, _p, _i64, _i32, _i32, _f, _p, _p, _p],
    "pm_swiglu_bwd": [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i32, _i32, _p, _p, _p],
    "pm_vq_bwd": [_p, _i64, _p, _p, _i32, _p, _i64, _p, _f, _i32, _p, _p, _p, _p],
    "pm_unpatchify8_bwd": [_p, _p, _p, _i32, _i32, _i32, _i32, _p],
}
# workspace-size queries (return int64 element counts): name -> number of int32 arguments
WORKSPACE_QUERIES = {
    "pm_wgrad_workspace_floats": 3,
    "pm_colsum_workspace_floats": 2,
    "pm_layernorm_bwd_workspace_floats": 2,
    "pm_swiglu_bwd_workspace_floats": 2,
}


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pm_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: rc={rc} ({msg})")
