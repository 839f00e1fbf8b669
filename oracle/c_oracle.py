"""ctypes loader of oracle/c/liboracle_scan.so (C restatement of the mask scan) — TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent / "c"
_LIB = _DIR / "liboracle_scan.so"
_lib = None


def available() -> bool:
    if _LIB.exists():
        return True
    try:
        subprocess.run(["make", "-s", "-C", str(_DIR)], check=True, capture_output=True)
    except (OSError, subprocess.CalledProcessError):
        return False
    return _LIB.exists()


def _load():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/c/liboracle_scan.so missing and could not be built")
        _lib = C.CDLL(str(_LIB))
        _lib.oracle_mask_scan.restype = None
        _lib.oracle_mask_scan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int64,
                                          C.c_int, C.c_void_p]
    return _lib


def mask_scan(mask: np.ndarray, id2slot: np.ndarray, num_slots: int) -> np.ndarray:
    """mask u32 [B,H,W] (or [H,W]); id2slot int32 [L] or [B,L] -> int32 [B,N,5]."""
    mask = np.ascontiguousarray(mask, dtype=np.uint32)
    if mask.ndim == 2:
        mask = mask[None]
    lut = np.ascontiguousarray(id2slot, dtype=np.int32)
    B, H, W = mask.shape
    stride = 0 if lut.ndim == 1 else lut.shape[1]
    out = np.empty((B, num_slots, 5), dtype=np.int32)
    _load().oracle_mask_scan(mask.ctypes.data, B, H, W, lut.ctypes.data, lut.shape[-1], stride, num_slots,
                             out.ctypes.data)
    return out
