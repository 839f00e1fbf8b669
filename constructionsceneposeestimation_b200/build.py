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

SOURCES = ["abi.cu", "mask_scan.cu", "project.cu", "keypoints.cu", "emit.cu", "pointcloud.cu", "depth_stats.cu", "depth_viz.cu", "format.cu", "yolo_text.cu", "text_format.cu",
           "label_json.cpp", "host_io.cpp"]

# -fmad=false: every f32/f64 operation rounds on its own, exactly like the numpy oracle's
# elementwise arithmetic, so projections and flags are reproducible bit for bit.
COMPILE_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",
    "-Xcompiler",
    "-fPIC",
]
LINK_FLAGS = ["--shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = COMPILE_FLAGS + LINK_FLAGS
OBJ_DIR = PKG_DIR / "build"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


_NVCC_VERSION = None


def _nvcc_version() -> str:
    global _NVCC_VERSION
    if _NVCC_VERSION is None:
        try:
            _NVCC_VERSION = subprocess.run([_nvcc(), "--version"], capture_output=True, text=True).stdout.strip()
        except (OSError, RuntimeError):
            _NVCC_VERSION = "unknown"
    return _NVCC_VERSION


def _headers():
    return sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted(INCLUDE.glob("*.h"))


def _digest(files, with_compiler: bool) -> str:
    h = hashlib.sha256()
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    if with_compiler:
        h.update(_nvcc_version().encode())
    return h.hexdigest()


def _source_digest(with_compiler: bool = False) -> str:
    """Digest of every file that goes into the library: the SOURCES list itself (.cu and .cpp), every
    header, and the flags.  The compiler version is part of the stamp only where a compiler exists
    (the GPU box runs the prebuilt library and may not have the same nvcc on PATH)."""
    return _digest([CSRC / s for s in SOURCES] + _headers(), with_compiler)


def is_fresh() -> bool:
    if not (LIB_PATH.exists() and STAMP_PATH.exists()):
        return False
    stamp = STAMP_PATH.read_text().split()
    return bool(stamp) and stamp[0] == _source_digest()


def _compile_one(src: str, verbose: bool):
    """One translation unit -> build/<name>.o; skipped when its own digest (source + headers + flags +
    compiler) is unchanged, so touching one kernel recompiles one file."""
    obj = OBJ_DIR / (src + ".o")
    tag = OBJ_DIR / (src + ".digest")
    want = _digest([CSRC / src] + _headers(), True)
    if not verbose and obj.exists() and tag.exists() and tag.read_text().strip() == want:
        return obj, ""
    cmd = [_nvcc(), *COMPILE_FLAGS, f"-I{INCLUDE}", f"-I{CSRC}"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-c", str(CSRC / src), "-o", str(obj)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    tag.write_text(want + "\n")
    return obj, proc.stdout + proc.stderr


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every kernel into one shared object; no-op when sources are unchanged.  Translation units
    are compiled in parallel (one nvcc per file) and linked with ``nvcc --shared``."""
    if not force and is_fresh():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    OBJ_DIR.mkdir(exist_ok=True)
    if force:
        for f in OBJ_DIR.glob("*.digest"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda s: _compile_one(s, verbose), SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [_nvcc(), *LINK_FLAGS, "-Xcompiler", "-fPIC"] + [str(o) for o, _ in results] + ["-o", str(LIB_PATH)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed ({proc.returncode}):\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    STAMP_PATH.write_text(_source_digest() + "\n" + _nvcc_version().splitlines()[-1] + "\n")
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
