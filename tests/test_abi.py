"""The C-ABI library builds, loads and exports exactly what include/cspe.h declares (no GPU needed)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared_functions():
    text = (ROOT / "include" / "cspe.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cspe_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared_functions()
    for must in ("cspe_mask_scan", "cspe_project_objects", "cspe_keypoints", "cspe_emit", "cspe_version",
                 "cspe_last_error", "cspe_depth_to_pointcloud", "cspe_depth_stats"):
        assert must in names


def test_library_exports_every_declared_symbol(libcspe_path):
    lib = ctypes.CDLL(str(libcspe_path))
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in cspe.h but not exported by libcspe.so"


def test_python_prototypes_cover_header(libcspe_path):
    from constructionsceneposeestimation_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == _declared_functions()
    lib = _lib.load()
    assert lib.cspe_version() == _lib.ABI_VERSION
    assert lib.cspe_last_error() is not None


def test_record_layouts_match_oracle():
    from constructionsceneposeestimation_b200 import _lib
    from oracle import labels as O

    assert _lib.RECORD_DTYPE == O.RECORD_DTYPE and _lib.RECORD_DTYPE.itemsize == 408
    assert _lib.BBOX3D_DTYPE == O.BBOX3D_DTYPE and _lib.BBOX3D_DTYPE.itemsize == 96
    assert (_lib.CAM_STRIDE, _lib.POSE_STRIDE, _lib.NUM_CLASSES) == (O.CAM_STRIDE, O.POSE_STRIDE, O.NUM_CLASSES)
    h = (ROOT / "include" / "cspe.h").read_text()
    assert f"#define CSPE_CAM_STRIDE {_lib.CAM_STRIDE}" in h and f"#define CSPE_POSE_STRIDE {_lib.POSE_STRIDE}" in h


def test_argument_validation_without_gpu(libcspe_path):
    """Bad arguments are rejected with a negative code and a message before any CUDA call."""
    from constructionsceneposeestimation_b200 import _lib

    lib = _lib.load()
    rc = lib.cspe_mask_scan(None, -1, 4, 4, None, 0, 0, 1, None, None)
    assert rc == -1 and b"negative" in lib.cspe_last_error()
    rc = lib.cspe_project_objects(None, 8, 1, None, None, 1, 1, None, None, None, None, None, None)
    assert rc == -1 and b"rec_stride" in lib.cspe_last_error()
    with pytest.raises(_lib.CspeError):
        _lib.check("cspe_project_objects", rc)
    assert lib.cspe_mask_scan(None, 0, 4, 4, None, 0, 0, 1, None, None) == 0  # empty batch is a no-op
    # round-2 entry points: object-level records, device text, row packing
    assert lib.cspe_union_records(None, 96, 4, 2, None, 0, None, 0, 1, 3, None) == -1      # base + U > recs_per_frame
    assert b"exceeds recs_per_frame" in lib.cspe_last_error()
    assert lib.cspe_union_records(None, 96, 8, 2, None, 0, None, 0, 1, 3, None) == -1 and b"null" in lib.cspe_last_error()
    assert lib.cspe_union_records(None, 96, 8, 2, None, 0, None, 0, 0, 3, None) == 0       # nothing to do
    assert lib.cspe_union_records(None, 90, 8, 2, None, 0, None, 0, 1, 3, None) == -1 and b"rec_stride" in lib.cspe_last_error()
    assert lib.cspe_format_coco(None, None, 1, 4, None, None, 64, None, None) == -1 and b"null" in lib.cspe_last_error()
    assert lib.cspe_format_coco(None, None, 0, 4, None, None, 64, None, None) == 0
    assert lib.cspe_format_yolo(None, None, -1, 4, None, 64, None, None) == -1 and b"negative" in lib.cspe_last_error()
    assert lib.cspe_pack_rows(None, 64, None, 2, None, 0, None, None) == -1 and b"null" in lib.cspe_last_error()
    assert lib.cspe_format_coco_images_host(0, 3, 64, 48, None, 10) == -1                   # capacity without a buffer


def test_no_cpu_fallback_in_product():
    """The product package never imports the oracle, and ops refuse CPU tensors."""
    import torch
    from constructionsceneposeestimation_b200 import ops

    pkg = ROOT / "constructionsceneposeestimation_b200"
    for py in pkg.glob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{py.name} imports the oracle"
    with pytest.raises(ValueError, match="CUDA"):
        ops.mask_scan(torch.zeros((1, 4, 4), dtype=torch.int32), torch.zeros(4, dtype=torch.int32), 2)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from constructionsceneposeestimation_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libcspe.so")
    with pytest.raises(_lib.CspeLibraryError):
        _lib.load()
