"""Build libcspe.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: ``python -m constructionsceneposeestimation_b200.build [--force] [--verbose]``.
The .so lands next to this file so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
INCLUDE = REPO_ROOT / "include"
LIB_PATH = PKG_DIR / "libcspe.so"
STAMP_PATH = PKG_DIR / ".libcspe.stamp"

SOURCES = ["abi.cu", "mask_scan.cu", "project.cu", "keypoints.cu", "emit.cu", "pointcloud.cu", "depth_stats.cu", "depth_viz.cu", "format.cu", "text_format.cu",
           "label_json.cpp"]

# -fmad=false: every f32/f64 operation rounds on its own, exactly like the numpy oracle's
# elementwise arithmetic, so projections and flags are reproducible bit for bit.
NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",
    "-Xcompiler",
    "-fPIC",
    "--shared",
    "-cudart",
    "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def _source_digest() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    return LIB_PATH.exists() and STAMP_PATH.exists() and STAMP_PATH.read_text().strip() == _source_digest()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every kernel into one shared object; no-op when sources are unchanged."""
    if not force and is_fresh():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, f"-I{INCLUDE}", f"-I{CSRC}"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(CSRC / s) for s in SOURCES] + ["-o", str(LIB_PATH)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        sys.stderr.write(proc.stdout + proc.stderr)
    STAMP_PATH.write_text(_source_digest() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
