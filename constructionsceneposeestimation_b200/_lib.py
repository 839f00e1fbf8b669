"""ctypes binding of libcspe.so — the C-ABI boundary declared in include/cspe.h.

The product path has NO CPU fallback: if the library is missing or fails to load, every
caller gets :class:`CspeLibraryError`.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import os

# CSPE_LIB points at an alternative build of the same ABI (used to A/B kernel variants on the GPU box)
LIB_PATH = Path(os.environ.get("CSPE_LIB") or Path(__file__).resolve().parent / "libcspe.so")

ABI_VERSION = 1
CAM_STRIDE = 24
POSE_STRIDE = 16
SCAN_FIELDS = 5
NUM_CLASSES = 10
BBOX3D_RECORD_BYTES = 96

OBJ_HAS_RECORD, OBJ_ANY_FRONT, OBJ_ALL_FRONT, OBJ_POSE_VALID, OBJ_APPROX_RECORD = 1, 2, 4, 8, 16
OBJ_RECORD_APPROX_BIT = 1 << 30   # OR into an obj_record entry: the record only approximates the object
KP_OUT, KP_OCCLUDED, KP_VISIBLE = 0, 1, 2

# host view of `cspe_record` (include/cspe.h); emit.cu static_asserts the 408-byte size
RECORD_DTYPE = np.dtype(
    [
        ("frame", "<i4"),
        ("inst_idx", "<i4"),
        ("class_id", "<i4"),
        ("count", "<i4"),
        ("x_min", "<i4"),
        ("y_min", "<i4"),
        ("x_max", "<i4"),
        ("y_max", "<i4"),
        ("flags", "<i4"),
        ("loose", "<i4", (4,)),
        ("pad0", "<i4"),
        ("occlusion", "<f4"),
        ("fill", "<f4"),
        ("truncation", "<f4"),
        ("visible_frac", "<f4"),
        ("yolo", "<f4", (4,)),
        ("uv", "<f8", (8, 2)),
        ("z", "<f8", (8,)),
        ("pose", "<f8", (POSE_STRIDE,)),
    ]
)
assert RECORD_DTYPE.itemsize == 408

# host view of `cspe_depth_stats_t`
DEPTH_STATS_DTYPE = np.dtype(
    [
        ("valid_pixels", "<i8"),
        ("zero_pixels", "<i8"),
        ("inf_pixels", "<i8"),
        ("total_pixels", "<i8"),
        ("depth_min", "<f4"),
        ("depth_max", "<f4"),
        ("depth_sum", "<f8"),
    ]
)
assert DEPTH_STATS_DTYPE.itemsize == 48

# Replicator bounding_box_3d record as the reference indexes it (gcd.py:562-564)
BBOX3D_DTYPE = np.dtype(
    [
        ("semanticId", "<u4"),
        ("x_min", "<f4"),
        ("y_min", "<f4"),
        ("z_min", "<f4"),
        ("x_max", "<f4"),
        ("y_max", "<f4"),
        ("z_max", "<f4"),
        ("transform", "<f4", (4, 4)),
        ("occlusionRatio", "<f4"),
    ]
)
assert BBOX3D_DTYPE.itemsize == BBOX3D_RECORD_BYTES


class CspeLibraryError(RuntimeError):
    """libcspe.so is missing, stale or failed to load — there is no fallback path."""


class CspeError(RuntimeError):
    """A libcspe entry point returned a negative status."""

    def __init__(self, func: str, code: int, message: str):
        super().__init__(f"{func} failed with {code}: {message}")
        self.func, self.code, self.message = func, code, message


_P = C.c_void_p
_I = C.c_int
_I64 = C.c_int64

# name -> (restype, argtypes); mirrors include/cspe.h one to one
PROTOTYPES = {
    "cspe_version": (_I, []),
    "cspe_last_error": (C.c_char_p, []),
    "cspe_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "cspe_mask_scan": (_I, [_P, _I, _I, _I, _P, _I, _I64, _I, _P, _P]),
    "cspe_mask_scan_accumulate": (_I, [_P, _I, _I, _I, _P, _I, _I64, _I, _P, _P]),
    "cspe_mask_scan_accumulate_overlapped": (_I, [_P, _I, _I, _I, _P, _I, _I64, _I, _P, _P]),
    "cspe_mask_scan_fills_device": (_I, [_I, _I, _I]),
    "cspe_mask_scan_depth_stats": (_I, [_P, _P, _I, _I, _I, _P, _I, _I64, _I, _P, _P, _P]),
    "cspe_project_objects": (_I, [_P, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "cspe_project_objects_overlapped": (_I, [_P, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "cspe_union_records": (_I, [_P, _I, _I, _I, _P, _I64, _P, _I64, _I, _I, _P]),
    "cspe_keypoints": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, C.c_double, _P, _P, _P, _P]),
    "cspe_keypoints_overlapped": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, C.c_double, _P, _P, _P, _P]),
    "cspe_emit": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "cspe_emit_reset_scan": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "cspe_emit_reset_scan_indirect": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "cspe_format_yolo": (_I, [_P, _P, _I, _I, _P, _I64, _P, _P]),
    "cspe_format_coco": (_I, [_P, _P, _I, _I, _P, _P, _I64, _P, _P]),
    "cspe_pack_rows": (_I, [_P, _I64, _P, _I, _P, _I64, _P, _P]),
    "cspe_memcpy_async": (_I, [_P, _P, C.c_size_t, _P]),
    "cspe_graph_edge_kinds": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "cspe_pointcloud_workspace_bytes": (C.c_size_t, [_I, _I]),
    "cspe_depth_to_pointcloud": (_I, [_P, _P, _I, _I, _I, _P, _P, _I64, _P, _P, _P]),
    "cspe_pointcloud_batch_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "cspe_depth_to_pointcloud_batch": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I64, _P, _P, _P]),
    "cspe_depth_stats": (_I, [_P, _I, _I, _I, _P, _P]),
    "cspe_depth_colormap": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "cspe_rgb_to_bgr": (_I, [_P, _I, _I64, _P, _P]),
    "cspe_text_workspace_bytes": (C.c_size_t, [_I64, _I]),
    "cspe_format_fixed6": (_I, [_P, _I, _I64, _P, _I, C.c_char_p, _P, _I64, _P, _I64, _P, _P, _P]),
    "cspe_write_files_host": (_I64, [C.c_char_p, C.c_char_p, _I, C.c_char_p, _I64, _I, _P, _I64, _P]),
    "cspe_format_coco_images_host": (_I64, [_I64, _I, _I, _I, _P, _I64]),
    "cspe_concat_rows_host": (_I64, [_P, _I64, _P, _I, _P, _I64]),
    "cspe_format_yolo_host": (_I64, [_P, _P, _I, _I, _I, _P, _I64, _P]),
    "cspe_format_coco_host": (_I64, [_P, _P, _I, _I, _I, _P, _I64, _P, _P, _P, _I, _I, _P, _I64]),
    "cspe_format_label_json_host": (_I64, [_P, _I, _I64, _P, C.c_char_p, C.c_char_p, _P, _P, _I, _I, _I, _P, _P, _P, _I,
                                           _I, _P, _I64]),
}

_lib = None


def load() -> C.CDLL:
    """Load libcspe.so once and attach prototypes; raises CspeLibraryError if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CspeLibraryError(
            f"{LIB_PATH} not found: build it with `python -m constructionsceneposeestimation_b200.build` "
            "(there is no CPU fallback)"
        )
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover - depends on the machine
        raise CspeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise CspeLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    if lib.cspe_version() != ABI_VERSION:
        raise CspeLibraryError(f"{LIB_PATH} has ABI {lib.cspe_version()}, expected {ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def check(func: str, code: int) -> None:
    if code < 0:
        msg = load().cspe_last_error()
        raise CspeError(func, code, msg.decode("utf-8", "replace") if msg else "")
