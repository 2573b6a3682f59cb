"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/freepose_b200.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (ROOT / "include" / "freepose_b200.h").read_text()
    return sorted(set(re.findall(r"FP_API\s+[\w\s\*]+?\b(fp_\w+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    syms = header_symbols()
    for must in ("fp_vit_forward", "fp_gemm_bf16", "fp_attention_bf16", "fp_score_topk", "fp_rasterize",
                 "fp_crop_resize_pad", "fp_depth_extents", "fp_ffa_pool", "fp_last_error"):
        assert must in syms
    assert len(syms) >= 19


def test_library_exports_every_declared_symbol(lib):
    from freepose_b200 import _lib
    syms = header_symbols()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"
    assert lib.fp_abi_version() == 3
    # no torch / C++ types leak through the boundary: only the declared C symbols are default-visible
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(syms), exported ^ set(syms)


def test_sass_contains_blackwell_tensor_and_tma_instructions(lib):
    from freepose_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in sass, "TMA loads missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    # legacy mma.sync (HMMA) is allowed in exactly one place: the <= 8 leftover query rows of the short-sequence attention
    # kernels (5 of 261 rows: < 2 % of the attention FLOPs); every GEMM and the full attention tiles are tcgen05
    per_fn = sass.split("Function : ")
    with_hmma = [blk.split("\n", 1)[0] for blk in per_fn if "HMMA." in blk.replace("UTCHMMA", "")]
    assert all("attention_kernel" in name or "attention_split_kernel" in name for name in with_hmma), with_hmma
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True,
                                       text=True).stdout


def test_argument_errors_are_reported_not_thrown(lib):
    """Contract violations return -1 with a message; nothing touches the (absent) GPU."""
    rc = lib.fp_ffa_pool(None, None, 1, 225, 1024, None, None, None)
    assert rc == -1 and b"multiple of 14" in lib.fp_last_error()
    n = ctypes.c_size_t(0)
    assert lib.fp_raster_workspace_bytes(2, 10, 16, 224, 3, ctypes.byref(n)) == -1
    assert b"msaa" in lib.fp_last_error()
    assert lib.fp_raster_workspace_bytes(2, 10, 16, 224, 4, ctypes.byref(n)) == 0 and n.value > 2 * 224 * 224 * 4 * 8
    assert lib.fp_vit_workspace_bytes(1024, 4096, 1, 224) > 261 * (1024 * 2 * 2 + 3072 * 2 + 4096 * 2)
    assert lib.fp_vit_workspace_bytes(768, 3072, 1, 518) > 1374 * (768 * 2 * 2 + 2304 * 2 + 3072 * 2)
