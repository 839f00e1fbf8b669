"""B200-native per-frame annotation hot path for construction-scene pose-estimation labels.

A from-scratch replacement of the label math behind ``generate_construction_data.py`` of
xander683/ConstructionScenePoseEstimation: instance-mask scan -> tight 2D boxes and pixel
counts, 3D-box corner projection, object-in-camera 6-DoF poses, skeleton keypoints with
depth-buffer visibility, occlusion ratios and COCO/YOLO/JSON records.  Host code is Python
with a Replicator-Writer-style ``write(data)`` surface; all pixel work runs in hand-written
sm_100a CUDA kernels behind the C ABI in ``include/cspe.h`` (``libcspe.so``).  There is no
CPU fallback: importing the writer without the built library raises.
"""
from . import _lib
from ._lib import CspeError, CspeLibraryError

__all__ = ["ConstructionLabelWriter", "CspeError", "CspeLibraryError", "_lib"]


def __getattr__(name):  # lazy: `import torch` is slow and not needed for the pure-host modules
    if name == "ConstructionLabelWriter":
        from .writer import ConstructionLabelWriter

        return ConstructionLabelWriter
    raise AttributeError(name)
