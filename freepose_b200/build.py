"""Builds ``libfreepose_b200.so`` (the C-ABI library) in-tree with nvcc for sm_100a.

``python -m freepose_b200.build`` or :func:`build`.  nvcc cross-compiles without a GPU.  The ``.so`` is
git-ignored but travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libfreepose_b200.so"

SOURCES = ["common.cu", "gemm.cu", "attention.cu", "attention_long.cu", "attention_split.cu", "attention_pair.cu", "elementwise.cu", "score.cu", "retrieval.cu", "raster.cu", "geometry.cu", "refiner.cu", "comm.cu",
           "vit.cu", "capi.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-I", str(PKG.parent / "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp(src: Path) -> str:
    h = hashlib.sha1()
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + sorted((PKG.parent / "include").glob("*.h")):
        h.update(hdr.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False, ptxas_info: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    nvcc = _nvcc()
    flags = list(NVCC_FLAGS) + (["-Xptxas", "-v"] if ptxas_info else [])
    jobs = []
    objs = []
    for name in SOURCES:
        src = CSRC / name
        if not src.exists():
            raise FileNotFoundError(src)
        obj = OBJ / (name + ".o")
        stamp_file = OBJ / (name + ".stamp")
        stamp = _stamp(src)
        objs.append(obj)
        if not force and not ptxas_info and obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
            continue
        jobs.append((name, [nvcc, *flags, "-c", str(src), "-o", str(obj)], stamp_file, stamp))

    def run(job):
        name, cmd, stamp_file, stamp = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
        if verbose or ptxas_info:
            sys.stderr.write(r.stdout + r.stderr)
        stamp_file.write_text(stamp)
        return name

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))

    if jobs or not LIB.exists():
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(path)
