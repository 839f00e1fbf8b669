"""Camera conventions of the reference, packed into the per-frame block the kernels read.

SURVEY §8a rows R4 and R5:
  * pose 7-vector ``[x, y, z, qx, qy, qz, qw]`` = camera position and camera->world rotation
    in USD camera axes (-Z forward, +Y up), gcd.py:587-605;
  * intrinsics from aperture / focal length, gcd.py:646-649 and 2035-2053.
"""
from __future__ import annotations

from collections.abc import Mapping
from typing import Optional, Sequence

import numpy as np

from ._lib import CAM_STRIDE

# the reference's fallback constants when the Camera API fails (gcd.py:2047-2053, 643-645)
DEFAULT_FOCAL_LENGTH = 18.14
DEFAULT_HORIZONTAL_APERTURE = 20.955
DEFAULT_VERTICAL_APERTURE = 15.2908
# clipping range set at gcd.py:1437
DEFAULT_NEAR, DEFAULT_FAR = 0.5, 250.0


def camera_params(width: int, height: int, focal_length: float = 12.0, horizontal_aperture: float = 25.0) -> dict:
    """The ``camera_params`` dict the reference writes (gcd.py:2036-2045); defaults are the
    script's own focal length / aperture (gcd.py:1442-1443)."""
    return {
        "horizontal_aperture": horizontal_aperture,
        "vertical_aperture": horizontal_aperture * (height / width),
        "focal_length": focal_length,
        "width": width,
        "height": height,
    }


def intrinsics(params: Mapping, width: Optional[int] = None, height: Optional[int] = None):
    """(fx, fy, cx, cy) exactly as gcd.py:639-649 derives them."""
    f = params.get("focal_length", DEFAULT_FOCAL_LENGTH)
    ha = params.get("horizontal_aperture", DEFAULT_HORIZONTAL_APERTURE)
    va = params.get("vertical_aperture", DEFAULT_VERTICAL_APERTURE)
    w = params.get("width", width)
    h = params.get("height", height)
    fx = (w * f) / ha
    fy = (h * f) / va
    return fx, fy, w / 2.0, h / 2.0


def quat_xyzw_to_matrix(q: Sequence[float]) -> np.ndarray:
    """Rotation matrix of a scalar-last quaternion, normalised first (what scipy's
    ``Rotation.from_quat(q).as_matrix()`` returns; used at gcd.py:681)."""
    q = np.asarray(q, dtype=np.float64)
    n = np.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    if not n > 0.0:
        raise ValueError("zero-norm camera quaternion")
    x, y, z, w = q / n
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    return np.array(
        [
            [x2 - y2 - z2 + w2, 2.0 * (xy - zw), 2.0 * (xz + yw)],
            [2.0 * (xy + zw), -x2 + y2 - z2 + w2, 2.0 * (yz - xw)],
            [2.0 * (xz - yw), 2.0 * (yz + xw), -x2 - y2 + z2 + w2],
        ],
        dtype=np.float64,
    )


def pose_from_usd_matrix(m: np.ndarray):
    """(t, Rcw) from a USD local-to-world 4x4 (row-vector convention): translation is the last
    row, and the column-convention rotation is the transpose of the upper 3x3 (gcd.py:599-601)."""
    m = np.asarray(m, dtype=np.float64).reshape(4, 4)
    return m[3, :3].copy(), m[:3, :3].T.copy()


def pack_camera(pose7: Sequence[float], params: Mapping, near: float = DEFAULT_NEAR, far: float = DEFAULT_FAR,
                out: Optional[np.ndarray] = None) -> np.ndarray:
    """One ``double[CAM_STRIDE]`` camera block (layout in include/cspe.h)."""
    blk = np.zeros(CAM_STRIDE, dtype=np.float64) if out is None else out
    pose7 = np.asarray(pose7, dtype=np.float64)
    blk[0:3] = pose7[0:3]
    blk[3:12] = quat_xyzw_to_matrix(pose7[3:7]).reshape(9)
    fx, fy, cx, cy = intrinsics(params)
    blk[12:16] = (fx, fy, cx, cy)
    blk[16], blk[17] = near, far
    blk[18], blk[19] = params["width"], params["height"]
    blk[20:24] = 0.0
    return blk


def pack_cameras(poses: Sequence[Sequence[float]], params: Sequence[Mapping], clips: Sequence[Sequence[float]],
                 out: Optional[np.ndarray] = None) -> np.ndarray:
    """``pack_camera`` for a batch in one set of array operations: ``double[B][CAM_STRIDE]``, bit-identical to B calls
    (the same IEEE operations per element, in the same order)."""
    B = len(poses)
    blk = np.zeros((B, CAM_STRIDE), dtype=np.float64) if out is None else out
    p = np.asarray(poses, dtype=np.float64).reshape(B, 7)
    q = p[:, 3:7]
    n = np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    if not np.all(n > 0.0):
        raise ValueError("zero-norm camera quaternion")
    x, y, z, w = (q / n[:, None]).T
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    blk[:, 0:3] = p[:, 0:3]
    blk[:, 3] = x2 - y2 - z2 + w2
    blk[:, 4] = 2.0 * (xy - zw)
    blk[:, 5] = 2.0 * (xz + yw)
    blk[:, 6] = 2.0 * (xy + zw)
    blk[:, 7] = -x2 + y2 - z2 + w2
    blk[:, 8] = 2.0 * (yz - xw)
    blk[:, 9] = 2.0 * (xz - yw)
    blk[:, 10] = 2.0 * (yz + xw)
    blk[:, 11] = -x2 - y2 + z2 + w2
    tail = np.empty((B, 8), dtype=np.float64)
    last_prm = last_clip = last_row = None
    for i, (prm, clip) in enumerate(zip(params, clips)):
        if last_row is None or not ((prm is last_prm or prm == last_prm) and clip == last_clip):   # cameras of a run rarely change
            fx, fy, cx, cy = intrinsics(prm)
            last_prm, last_clip = prm, clip
            last_row = (fx, fy, cx, cy, clip[0], clip[1], prm["width"], prm["height"])
        tail[i] = last_row
    blk[:, 12:20] = tail
    blk[:, 20:24] = 0.0
    return blk


def matrix_to_quat_xyzw(r: np.ndarray) -> np.ndarray:
    """Scalar-last quaternion of a rotation matrix, the way scipy's ``Rotation.from_matrix(r).as_quat()``
    produces it at gcd.py:602 (largest of the three diagonal entries and the trace picks the branch; the sign
    is whatever that branch gives — not canonicalised, like the reference's)."""
    m = np.asarray(r, dtype=np.float64).reshape(3, 3)
    decision = np.array([m[0, 0], m[1, 1], m[2, 2], m[0, 0] + m[1, 1] + m[2, 2]])
    choice = int(np.argmax(decision))
    q = np.empty(4, dtype=np.float64)
    if choice != 3:
        i = choice
        j = (i + 1) % 3
        k = (j + 1) % 3
        q[i] = 1.0 - decision[3] + 2.0 * m[i, i]
        q[j] = m[j, i] + m[i, j]
        q[k] = m[k, i] + m[i, k]
        q[3] = m[k, j] - m[j, k]
    else:
        q[0] = m[2, 1] - m[1, 2]
        q[1] = m[0, 2] - m[2, 0]
        q[2] = m[1, 0] - m[0, 1]
        q[3] = 1.0 + decision[3]
    return q / np.sqrt(q @ q)


def is_replicator_camera_params(params) -> bool:
    """True for the dict Replicator's ``camera_params`` annotator delivers (``cameraViewTransform``,
    ``cameraFocalLength``, ``cameraAperture``, ``renderProductResolution`` ...) as opposed to the reference's own
    five-field dict (gcd.py:2039-2045)."""
    return isinstance(params, Mapping) and "cameraViewTransform" in params


def from_replicator_camera_params(params: Mapping, width: Optional[int] = None, height: Optional[int] = None):
    """(pose7, reference-style params dict) from Replicator's native ``camera_params`` payload.

    ``cameraViewTransform`` is the world -> camera matrix in USD's row-vector convention (16 values, row-major):
    its inverse is the camera prim's local-to-world matrix, from which the pose follows exactly as ``get_obj_pose``
    takes it (gcd.py:596-605): translation = last row, rotation = transposed upper 3x3, scipy quaternion.
    Intrinsics follow gcd.py:2036-2045 / 646-649: ``focal_length`` = cameraFocalLength, ``horizontal_aperture`` =
    cameraAperture[0], ``vertical_aperture`` = horizontal * H / W, W x H = renderProductResolution (units cancel
    in fx = W * f / aperture)."""
    view = np.asarray(params["cameraViewTransform"], dtype=np.float64).reshape(4, 4)
    t, rcw = pose_from_usd_matrix(np.linalg.inv(view))
    pose7 = np.concatenate([t, matrix_to_quat_xyzw(rcw)])
    res = params.get("renderProductResolution")
    if res is not None and len(res) >= 2:
        width, height = int(res[0]), int(res[1])
    if width is None or height is None:
        raise ValueError("camera_params has no renderProductResolution and no image shape is known")
    aperture = params.get("cameraAperture")
    ha = float(np.asarray(aperture).reshape(-1)[0]) if aperture is not None else DEFAULT_HORIZONTAL_APERTURE
    f = params.get("cameraFocalLength")
    f = float(np.asarray(f).reshape(-1)[0]) if f is not None else DEFAULT_FOCAL_LENGTH
    out = {"horizontal_aperture": ha, "vertical_aperture": ha * (height / width), "focal_length": f,
           "width": width, "height": height}
    near_far = params.get("cameraNearFar")
    clip = (float(near_far[0]), float(near_far[1])) if near_far is not None and len(near_far) >= 2 else None
    return [float(v) for v in pose7], out, clip
