"""The C-ABI library loads and exports every symbol include/paintmind_b200.h declares (no GPU needed)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "paintmind_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t|const char\*)\s+(pm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("pm_gemm_bf16", "pm_attn_fwd", "pm_vq_fwd", "pm_vq_codebook_prep", "pm_vq_gather", "pm_layernorm",
                 "pm_patchify8", "pm_maskgit_sample", "pm_maskgit_remask", "pm_split_rows32", "pm_cast_f32_bf16",
                 # generator backward path
                 "pm_attn_bwd", "pm_wgrad_bf16", "pm_wgrad_workspace_floats", "pm_colsum_bf16", "pm_layernorm_bwd",
                 "pm_swiglu_bwd", "pm_vq_bwd", "pm_unpatchify8_bwd"):
        assert must in syms


def test_library_builds_loads_and_exports_all_declared_symbols():
    from paintmind_b200 import _lib, build
    lib_path = build.build()
    assert lib_path.exists()
    lib = ctypes.CDLL(str(lib_path))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    # the Python binding table covers every compute entry point of the header
    bound = set(_lib.EXPORTS) | set(_lib.WORKSPACE_QUERIES) | {"pm_version", "pm_device_check", "pm_error_string"}
    assert set(declared_symbols()) <= bound
    loaded = _lib.load()
    assert loaded.pm_version() == 1
    assert loaded.pm_error_string(-1).decode().startswith("invalid argument")


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the argument structs: natural alignment, pointer/int64 fields first then int32."""
    from paintmind_b200 import _lib
    assert ctypes.sizeof(_lib.GemmArgs) == 8 * 8 + 5 * 8 + 12 * 4 + 4 + 4 + 8 + 8 + 8 + 8  # 8 ptrs, 5 i64, 12 i32, f32, pad, ptr, i32+pad, ptr, i32+pad
    assert ctypes.sizeof(_lib.AttnArgs) == 4 * 8 + 8 * 8 + 5 * 4 + 4 + 4 * 8 + 2 * 4  # + lse, lse_ld, o32, ldo32, q_prescaled, reserved
    assert ctypes.sizeof(_lib.AttnBwdArgs) == 10 * 8 + 16 * 8 + 5 * 4 + 4 + 4 + 4 + 2 * 8   # 10 ptrs, 16 i64, 5 i32, f32, i32, pad, i64, ptr
    assert ctypes.sizeof(_lib.VqArgs) == 10 * 8 + 8 + 4 * 4
    assert ctypes.sizeof(_lib.MaskgitSampleArgs) == 5 * 8 + 3 * 8 + 2 * 8 + 3 * 4 + 4 + 2 * 8      # + step_tab, step_idx


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """Evidence that the shipped kernels are tcgen05 / TMA code (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
    import shutil
    import subprocess
    from paintmind_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-sass", str(build.build())], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out or "UTCMMA" in out, "no tcgen05.mma (UTC*MMA) in SASS"
    assert "LDTM" in out and "UTMALDG" in out and "UTMASTG" in out
    assert "HMMA.16816" not in out, "legacy mma.sync path present"
