"""numpy oracle of the annotation hot path — TEST INFRASTRUCTURE (see oracle/__init__.py).

Every function cites the span of ``/root/reference/generate_construction_data.py``
(``gcd.py``) it restates, or — for [SPEC] stages the reference does not implement — the
convention it follows and the SURVEY §8a row that defines it.  Float stages are written
as explicit elementwise IEEE operations in a fixed order (no BLAS, no FMA), which is the
order the CUDA kernels use with -fmad=false.
"""
from __future__ import annotations

import warnings
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
from scipy import ndimage
from scipy.spatial.transform import Rotation

CAM_STRIDE = 24
POSE_STRIDE = 16
NUM_CLASSES = 10
OBJ_HAS_RECORD, OBJ_ANY_FRONT, OBJ_ALL_FRONT, OBJ_POSE_VALID, OBJ_APPROX_RECORD = 1, 2, 4, 8, 16
OBJ_RECORD_APPROX_BIT = 1 << 30   # marks an obj_record entry whose record only approximates the object
KP_OUT, KP_OCCLUDED, KP_VISIBLE = 0, 1, 2

# Replicator bounding_box_3d record; field order is what gcd.py:562-564 indexes (1..6 extents,
# 7 transform)
BBOX3D_DTYPE = np.dtype(
    [("semanticId", "<u4"), ("x_min", "<f4"), ("y_min", "<f4"), ("z_min", "<f4"), ("x_max", "<f4"),
     ("y_max", "<f4"), ("z_max", "<f4"), ("transform", "<f4", (4, 4)), ("occlusionRatio", "<f4")]
)

RECORD_DTYPE = np.dtype(
    [("frame", "<i4"), ("inst_idx", "<i4"), ("class_id", "<i4"), ("count", "<i4"), ("x_min", "<i4"),
     ("y_min", "<i4"), ("x_max", "<i4"), ("y_max", "<i4"), ("flags", "<i4"), ("loose", "<i4", (4,)),
     ("pad0", "<i4"), ("occlusion", "<f4"), ("fill", "<f4"), ("truncation", "<f4"), ("visible_frac", "<f4"),
     ("yolo", "<f4", (4,)), ("uv", "<f8", (8, 2)), ("z", "<f8", (8,)), ("pose", "<f8", (POSE_STRIDE,))]
)


# ------------------------------------------------------------------------------------------
# R4 / R5: camera
# ------------------------------------------------------------------------------------------
def intrinsics(params: Mapping, w: Optional[int] = None, h: Optional[int] = None):
    """gcd.py:639-649: fx = W*f/ha, fy = H*f/va, cx = W/2, cy = H/2 (with the script's defaults)."""
    focal_length = params.get("focal_length", 18.14)
    horizontal_aperture = params.get("horizontal_aperture", 20.955)
    vertical_aperture = params.get("vertical_aperture", 15.2908)
    img_width = params.get("width", w)
    img_height = params.get("height", h)
    return ((img_width * focal_length) / horizontal_aperture, (img_height * focal_length) / vertical_aperture,
            img_width / 2.0, img_height / 2.0)


def camera_pose_from_usd_matrix(m) -> list:
    """gcd.py:587-605 without pxr (PARITY UNPINNED): GfMatrix4d is row-vector, so
    ExtractTranslation() is the last row and ExtractRotationMatrix().GetTranspose() is the
    column-convention camera->world rotation; scipy gives the scalar-last quaternion."""
    m = np.asarray(m, dtype=np.float64).reshape(4, 4)
    quat = Rotation.from_matrix(m[:3, :3].T).as_quat()
    return [m[3, 0], m[3, 1], m[3, 2], quat[0], quat[1], quat[2], quat[3]]


def pack_camera(pose7: Sequence[float], params: Mapping, near: float = 0.5, far: float = 250.0) -> np.ndarray:
    """Camera block double[24] (include/cspe.h).  Rotation via scipy exactly as gcd.py:681 does."""
    blk = np.zeros(CAM_STRIDE, dtype=np.float64)
    blk[0:3] = np.asarray(pose7[:3], dtype=np.float64)
    blk[3:12] = Rotation.from_quat(np.asarray(pose7[3:7], dtype=np.float64)).as_matrix().reshape(9)
    blk[12:16] = intrinsics(params)
    blk[16], blk[17] = near, far
    blk[18], blk[19] = params["width"], params["height"]
    return blk


# ------------------------------------------------------------------------------------------
# S1: instance-ID mask scan ([SPEC]; plugs the hole at gcd.py:1908-1910)
# ------------------------------------------------------------------------------------------
def _slot_image(mask: np.ndarray, id2slot: np.ndarray, num_slots: int) -> np.ndarray:
    mask = np.asarray(mask)
    if mask.dtype != np.uint32:
        mask = mask.astype(np.int64).astype(np.uint32) if mask.dtype.kind == "i" else mask.astype(np.uint32)
    lut = np.asarray(id2slot, dtype=np.int64)
    in_lut = mask < lut.shape[0]
    slots = np.full(mask.shape, -1, dtype=np.int64)
    if lut.shape[0]:
        slots[in_lut] = lut[mask[in_lut]]
    slots[(slots < 0) | (slots >= num_slots)] = -1
    return slots


def mask_scan_naive(mask: np.ndarray, id2slot: np.ndarray, num_slots: int) -> np.ndarray:
    """Per slot: count, x_min, y_min, x_max, y_max (inclusive) via ``np.nonzero(slot_img == n)``.
    Absent slot -> (0, W, H, -1, -1).  One frame [H,W] -> int32 [N,5]."""
    H, W = mask.shape
    slots = _slot_image(mask, id2slot, num_slots)
    out = np.empty((num_slots, 5), dtype=np.int32)
    for n in range(num_slots):
        ys, xs = np.nonzero(slots == n)
        if ys.size == 0:
            out[n] = (0, W, H, -1, -1)
        else:
            out[n] = (ys.size, xs.min(), ys.min(), xs.max(), ys.max())
    return out


def mask_scan_fast(mask: np.ndarray, id2slot: np.ndarray, num_slots: int) -> np.ndarray:
    """Same result from ``np.bincount`` + ``scipy.ndimage.find_objects`` (SURVEY §6 candidate B)."""
    H, W = mask.shape
    slots = _slot_image(mask, id2slot, num_slots)
    labels = (slots + 1).astype(np.int32)
    counts = np.bincount(labels.ravel(), minlength=num_slots + 1)[1:num_slots + 1]
    out = np.empty((num_slots, 5), dtype=np.int32)
    out[:, 0] = counts
    out[:, 1], out[:, 2], out[:, 3], out[:, 4] = W, H, -1, -1
    for n, sl in enumerate(ndimage.find_objects(labels, max_label=num_slots)):
        if sl is not None:
            out[n, 1:] = (sl[1].start, sl[0].start, sl[1].stop - 1, sl[0].stop - 1)
    return out


def mask_scan(mask: np.ndarray, id2slot: np.ndarray, num_slots: int, fast: bool = True) -> np.ndarray:
    """Batched: mask [B,H,W], id2slot [L] or [B,L] -> int32 [B,N,5]."""
    mask = np.asarray(mask)
    if mask.ndim == 2:
        mask = mask[None]
    id2slot = np.asarray(id2slot)
    fn = mask_scan_fast if fast else mask_scan_naive
    out = np.empty((mask.shape[0], num_slots, 5), dtype=np.int32)
    for b in range(mask.shape[0]):
        out[b] = fn(mask[b], id2slot if id2slot.ndim == 1 else id2slot[b], num_slots)
    return out


# ------------------------------------------------------------------------------------------
# R3: bbox3d record -> world centre / size / Euler  (gcd.py:553-584)
# ------------------------------------------------------------------------------------------
def bbox_to_transform(rec) -> Tuple[list, list, list]:
    """Restatement of bboxDict_to_transform.  ``rec`` indexes like the Replicator record:
    [1..6] local extents (f32), [7] 4x4 f32 transform in USD row-vector convention."""
    lo = np.array([rec[1], rec[2], rec[3]])          # gcd.py:562-563 (float32 array)
    hi = np.array([rec[4], rec[5], rec[6]])
    corner = np.array([lo, hi])
    m = rec[7].reshape(4, 4).T                        # gcd.py:568: column-vector form
    center_local = np.mean(corner, axis=0)            # f32 mean
    center_world = (m @ np.append(center_local, 1.0))[:3].tolist()   # promoted to f64
    rot = m[:3, :3]
    u, _, vt = np.linalg.svd(rot)                     # f32 LAPACK, like the reference
    euler = Rotation.from_matrix(np.dot(u, vt)).as_euler("xyz", degrees=True)  # raises if det <= 0
    scale = np.array([np.linalg.norm(rot[:, 0]), np.linalg.norm(rot[:, 1]), np.linalg.norm(rot[:, 2])])
    size_world = scale * np.abs(corner[1] - corner[0]).tolist()
    return center_world, size_world.tolist(), euler.tolist()


# ------------------------------------------------------------------------------------------
# S2 + S3: corner projection and object-in-camera pose ([SPEC])
# ------------------------------------------------------------------------------------------
def _cam_parts(cam: np.ndarray):
    t = cam[0:3]
    rcw = cam[3:12].reshape(3, 3)
    return t, rcw, cam[12], cam[13], cam[14], cam[15], cam[16]


def _to_camera(px, py, pz, t, rcw):
    """p_c = Rcw^T (p_w - t), elementwise in the kernels' operation order."""
    d0, d1, d2 = px - t[0], py - t[1], pz - t[2]
    return tuple((rcw[0, i] * d0 + rcw[1, i] * d1) + rcw[2, i] * d2 for i in range(3))


def quat_xyzw_canonical(r: np.ndarray) -> np.ndarray:
    """scipy's scalar-last quaternion of a rotation matrix, sign fixed to w >= 0."""
    q = Rotation.from_matrix(r).as_quat()
    return -q if q[3] < 0 else q


def polar_rotation(rot: np.ndarray) -> np.ndarray:
    """Orthogonal polar factor U @ Vt (gcd.py:573-574) in float64."""
    u, _, vt = np.linalg.svd(rot.astype(np.float64))
    return u @ vt


def union_records(records: np.ndarray, base: int, offsets: np.ndarray, members: np.ndarray) -> np.ndarray:
    """Object-level records for multi-mesh objects without a record of their own: the world-axis-aligned range
    the reference reads from the USD stage (BBoxCache.ComputeWorldBound(prim).ComputeAlignedRange(), gcd.py:2000-2009),
    restated as the min / max over the eight world-space corners (p_w = M @ [c,1], gcd.py:568) of every mesh record.
    records BBOX3D_DTYPE [B,R]; object u of frame f unites records members[f, offsets[f,u]:offsets[f,u+1]] (indices
    < base) and is written to records[f, base + u]: extents = the range (float32), transform = identity, semanticId /
    occlusionRatio of its first member.  offsets / members may be 1-D (shared by the batch).  Returns a copy."""
    out = np.array(records, copy=True)
    B = out.shape[0]
    offsets, members = np.asarray(offsets), np.asarray(members)
    U = offsets.shape[-1] - 1
    with np.errstate(all="ignore"):
        for f in range(B):
            off = offsets[f] if offsets.ndim == 2 else offsets
            mem = members[f] if members.ndim == 2 else members
            for u in range(U):
                mn = np.full(3, np.inf)
                mx = np.full(3, -np.inf)
                sem, occ, first = 0, np.float32(0), True
                for ri in mem[int(off[u]):int(off[u + 1])]:
                    ri = int(ri)
                    if ri < 0 or ri >= base:
                        continue
                    rec = out[f, ri]
                    if first:
                        sem, occ, first = rec["semanticId"], rec["occlusionRatio"], False
                    lo = np.array([rec["x_min"], rec["y_min"], rec["z_min"]], dtype=np.float32)
                    hi = np.array([rec["x_max"], rec["y_max"], rec["z_max"]], dtype=np.float32)
                    T = np.asarray(rec["transform"], dtype=np.float32).reshape(4, 4).astype(np.float64)
                    k = np.arange(8)
                    c0 = np.where(k & 1, hi[0], lo[0]).astype(np.float64)
                    c1 = np.where(k & 2, hi[1], lo[1]).astype(np.float64)
                    c2 = np.where(k & 4, hi[2], lo[2]).astype(np.float64)
                    for j in range(3):
                        pw = ((c0 * T[0, j] + c1 * T[1, j]) + c2 * T[2, j]) + T[3, j]
                        mn[j] = np.fmin(mn[j], np.fmin.reduce(pw))
                        mx[j] = np.fmax(mx[j], np.fmax.reduce(pw))
                new = np.zeros((), dtype=out.dtype)
                new["semanticId"] = sem
                new["occlusionRatio"] = occ
                ok = bool((mn <= mx).all())
                lo32 = mn.astype(np.float32) if ok else np.full(3, np.nan, dtype=np.float32)
                hi32 = mx.astype(np.float32) if ok else np.full(3, np.nan, dtype=np.float32)
                new["x_min"], new["y_min"], new["z_min"] = lo32
                new["x_max"], new["y_max"], new["z_max"] = hi32
                new["transform"] = np.eye(4, dtype=np.float32).reshape(new["transform"].shape)
                out[f, base + u] = new
    return out


def project_objects(records: np.ndarray, obj_record: np.ndarray, cam: np.ndarray):
    """records: BBOX3D_DTYPE [B,R]; obj_record int32 [B,N] (-1 = no record); cam f64 [B,24].
    Returns uv [B,N,8,2], z [B,N,8], pose [B,N,16], loose [B,N,4] (f64) and flags u8 [B,N].

    Corner k picks max where bit is set (bit0 x, bit1 y, bit2 z); p_w = M @ [c,1] with
    M = transform.reshape(4,4).T (gcd.py:568); pinhole in USD camera axes (gcd.py:587-605):
    z = -p_c.z, u = cx + fx*p_c.x/z, v = cy - fy*p_c.y/z; in_front = z > near."""
    records = np.asarray(records)
    B, N = obj_record.shape
    R = records.shape[1] if records.ndim == 2 else 0
    uv = np.full((B, N, 8, 2), np.nan)
    z = np.full((B, N, 8), np.nan)
    pose = np.full((B, N, POSE_STRIDE), np.nan)
    loose = np.full((B, N, 4), np.nan)
    flags = np.zeros((B, N), dtype=np.uint8)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in range(B):
            t, rcw, fx, fy, cx, cy, near = _cam_parts(cam[f])
            for n in range(N):
                ri = int(obj_record[f, n])
                approx = ri >= 0 and bool(ri & OBJ_RECORD_APPROX_BIT)
                if ri >= 0:
                    ri &= ~OBJ_RECORD_APPROX_BIT
                if ri < 0 or ri >= R:
                    continue
                rec = records[f, ri]
                lo = np.array([rec["x_min"], rec["y_min"], rec["z_min"]], dtype=np.float32)
                hi = np.array([rec["x_max"], rec["y_max"], rec["z_max"]], dtype=np.float32)
                T32 = np.asarray(rec["transform"], dtype=np.float32).reshape(4, 4)
                T = T32.astype(np.float64)
                k = np.arange(8)
                c0 = np.where(k & 1, hi[0], lo[0]).astype(np.float64)
                c1 = np.where(k & 2, hi[1], lo[1]).astype(np.float64)
                c2 = np.where(k & 4, hi[2], lo[2]).astype(np.float64)
                pw = [((c0 * T[0, j] + c1 * T[1, j]) + c2 * T[2, j]) + T[3, j] for j in range(3)]
                pc = _to_camera(pw[0], pw[1], pw[2], t, rcw)
                zc = -pc[2]
                u = cx + (fx * pc[0]) / zc
                v = cy - (fy * pc[1]) / zc
                front = zc > near
                uv[f, n, :, 0], uv[f, n, :, 1], z[f, n] = u, v, zc
                fl = OBJ_HAS_RECORD
                if front.any():
                    fl |= OBJ_ANY_FRONT
                    loose[f, n] = (u[front].min(), v[front].min(), u[front].max(), v[front].max())
                if front.all():
                    fl |= OBJ_ALL_FRONT
                # ---- pose ----
                cl = ((lo + hi) * np.float32(0.5)).astype(np.float64)          # f32 mean, gcd.py:566
                cw = np.array([((cl[0] * T[0, j] + cl[1] * T[1, j]) + cl[2] * T[2, j]) + T[3, j] for j in range(3)])
                n2 = (T32[:3, 0] * T32[:3, 0] + T32[:3, 1] * T32[:3, 1]) + T32[:3, 2] * T32[:3, 2]   # f32, rows of T
                size = np.sqrt(n2).astype(np.float64) * np.abs(hi - lo).astype(np.float64)
                rot = T[:3, :3].T                                              # gcd.py:568,572
                det = np.linalg.det(rot)
                ok = bool(np.isfinite(rot).all() and det > 0 and np.isfinite(cw).all())
                pc_c = _to_camera(cw[0], cw[1], cw[2], t, rcw)
                pose[f, n, 0:3] = pc_c
                pose[f, n, 7:10] = cw
                pose[f, n, 10:13] = size
                if ok:
                    rwo = polar_rotation(rot)
                    pose[f, n, 3:7] = quat_xyzw_canonical(rcw.T @ rwo)
                    pose[f, n, 13:16] = Rotation.from_matrix(rwo).as_euler("xyz", degrees=True)
                    fl |= OBJ_POSE_VALID
                if approx:
                    fl |= OBJ_APPROX_RECORD
                flags[f, n] = fl
    return uv, z, pose, loose, flags


# ------------------------------------------------------------------------------------------
# S4: skeleton keypoints + depth-buffer visibility ([SPEC])
# ------------------------------------------------------------------------------------------
def keypoints(joints: np.ndarray, depth: np.ndarray, cam: np.ndarray, tol: float = 0.15):
    """joints f32 [B,P,J,3] world; depth f32 [B,H,W] (inf = no hit, gcd.py:318-321); cam [B,24].
    -> kp f64 [B,P,J,2], kz f64 [B,P,J], vis u8 [B,P,J] (COCO 0 out / 1 occluded / 2 visible).
    in_view = 0 <= floor(u) < W and 0 <= floor(v) < H and z > near;
    visible = in_view and isfinite(d) and z <= d + tol with d = depth[floor(v), floor(u)]."""
    joints = np.asarray(joints, dtype=np.float32)
    B, P, J, _ = joints.shape
    H, W = depth.shape[1], depth.shape[2]
    kp = np.empty((B, P, J, 2))
    kz = np.empty((B, P, J))
    vis = np.zeros((B, P, J), dtype=np.uint8)
    with np.errstate(all="ignore"):
        for f in range(B):
            t, rcw, fx, fy, cx, cy, near = _cam_parts(cam[f])
            jw = joints[f].astype(np.float64)
            pc = _to_camera(jw[..., 0], jw[..., 1], jw[..., 2], t, rcw)
            zc = -pc[2]
            u = cx + (fx * pc[0]) / zc
            v = cy - (fy * pc[1]) / zc
            kp[f, ..., 0], kp[f, ..., 1], kz[f] = u, v, zc
            in_view = (u >= 0.0) & (u < float(W)) & (v >= 0.0) & (v < float(H)) & (zc > near)
            ui = np.where(in_view, np.floor(u), 0).astype(np.int64)
            vi = np.where(in_view, np.floor(v), 0).astype(np.int64)
            d = depth[f][vi, ui].astype(np.float64)
            visible = in_view & np.isfinite(d) & (zc <= d + tol)
            vis[f] = np.where(in_view, np.where(visible, KP_VISIBLE, KP_OCCLUDED), KP_OUT)
    return kp, kz, vis


# ------------------------------------------------------------------------------------------
# S5 + S6 + S7: ratios, stable compaction, records, class histogram ([SPEC])
# ------------------------------------------------------------------------------------------
def emit(scan, uv, z, pose, loose, flags, slot_class, H: int, W: int, min_pixels: int = 1, frame_base: int = 0):
    """-> (records RECORD_DTYPE [B,N] with the first n_out[f] rows of frame f valid and the
    rest zero, n_out int32 [B], class_hist int64 [10]).  Keep a slot iff class >= 0, count >=
    min_pixels and some corner is in front; order = increasing inst_idx (gcd.py:1876-1886).
    Ratios are float32: fill = count/tight_area; visible_frac = min(1, count/loose_area) with
    the loose box = projected 3D box clipped to the image and integerised
    (lx0 = max(0, floor(u_min)), lx1 = min(W-1, ceil(u_max)-1)); occlusion = 1 - visible_frac;
    truncation = 1 - clipped_area/unclipped_area of the continuous projected box."""
    B, N = slot_class.shape
    recs = np.zeros((B, N), dtype=RECORD_DTYPE)
    n_out = np.zeros(B, dtype=np.int32)
    hist = np.zeros(NUM_CLASSES, dtype=np.int64)
    f32 = np.float32
    with np.errstate(all="ignore"):
        for f in range(B):
            k = 0
            for n in range(N):
                cls, cnt, fl = int(slot_class[f, n]), int(scan[f, n, 0]), int(flags[f, n])
                if not (cls >= 0 and cnt >= min_pixels and (fl & OBJ_ANY_FRONT)):
                    continue
                r = recs[f, k]
                k += 1
                x0, y0, x1, y1 = (int(v) for v in scan[f, n, 1:5])
                r["frame"], r["inst_idx"], r["class_id"], r["count"] = frame_base + f, n, cls, cnt
                r["x_min"], r["y_min"], r["x_max"], r["y_max"], r["flags"] = x0, y0, x1, y1, fl
                tw, th = x1 - x0 + 1, y1 - y0 + 1
                umin, vmin, umax, vmax = (float(v) for v in loose[f, n])
                lx0 = int(min(max(np.floor(umin), 0.0), float(W)))
                ly0 = int(min(max(np.floor(vmin), 0.0), float(H)))
                lx1 = int(max(min(np.ceil(umax) - 1.0, W - 1.0), -1.0))
                ly1 = int(max(min(np.ceil(vmax) - 1.0, H - 1.0), -1.0))
                lw, lh = max(0, lx1 - lx0 + 1), max(0, ly1 - ly0 + 1)
                loose_area = lw * lh
                r["loose"] = (lx0, ly0, lx1, ly1) if loose_area > 0 else (0, 0, -1, -1)
                vis = np.minimum(f32(1.0), f32(cnt) / f32(loose_area)) if loose_area > 0 else f32(0.0)
                r["visible_frac"] = vis
                r["occlusion"] = f32(1.0) - vis
                r["fill"] = f32(cnt) / f32(tw * th) if cnt > 0 else f32(0.0)
                ua = (umax - umin) * (vmax - vmin)
                cwid = max(min(umax, float(W)) - max(umin, 0.0), 0.0)
                chei = max(min(vmax, float(H)) - max(vmin, 0.0), 0.0)
                ca = cwid * chei
                r["truncation"] = f32(1.0) - f32(ca) / f32(ua) if ua > 0.0 else f32(1.0)
                if cnt > 0:
                    r["yolo"] = ((f32(x0 + x1 + 1) * f32(0.5)) / f32(W), (f32(y0 + y1 + 1) * f32(0.5)) / f32(H),
                                 f32(tw) / f32(W), f32(th) / f32(H))
                r["uv"], r["z"], r["pose"] = uv[f, n], z[f, n], pose[f, n]
                if cls < NUM_CLASSES:
                    hist[cls] += 1
            n_out[f] = k
    return recs, n_out, hist


# ------------------------------------------------------------------------------------------
# f1: depth -> coloured point cloud  (gcd.py:616-711)
# ------------------------------------------------------------------------------------------
def depth_to_pointcloud(depth_data, rgb_image, camera_params: Mapping, camera_pose) -> Optional[np.ndarray]:
    """Restatement of depth_to_pointcloud_with_rgb: (N,6) float64 [x,y,z,r,g,b] in row-major
    pixel order, or None when no pixel is valid."""
    h, w = depth_data.shape
    fx, fy, cx, cy = intrinsics(camera_params, w, h)
    u, v = np.meshgrid(np.arange(w), np.arange(h))
    valid = np.isfinite(depth_data) & (depth_data > 0) & (depth_data < 250)      # gcd.py:655
    if not valid.any():
        return None
    zc = depth_data[valid]
    xc = (u[valid] - cx) * zc / fx                                               # gcd.py:664-666
    yc = (v[valid] - cy) * zc / fy
    pts = np.stack([xc, yc, zc], axis=-1)
    rot = Rotation.from_quat(np.array(camera_pose[3:])).as_matrix()               # gcd.py:677-681
    world = (rot @ pts.T).T + np.array(camera_pose[:3])                           # gcd.py:685
    if rgb_image is not None and rgb_image.size > 0:
        if rgb_image.shape[2] >= 3:
            rgb = rgb_image[valid, :3]
            rgb = (rgb * 255).astype(np.uint8) if rgb.max() <= 1.0 else rgb.astype(np.uint8)   # gcd.py:693-696
        else:
            rgb = np.ones((world.shape[0], 3), dtype=np.uint8) * 255
    else:
        rgb = np.ones((world.shape[0], 3), dtype=np.uint8) * 255
    return np.hstack([world, rgb])


# ------------------------------------------------------------------------------------------
# f2: depth statistics  (gcd.py:314-359)
# ------------------------------------------------------------------------------------------
def depth_stats(depth_data: np.ndarray) -> Dict[str, object]:
    """Restatement of the numeric part of DataQualityLogger.log_depth: the dict it stores under
    current_frame['depth'] (gcd.py:333-342)."""
    ok = np.isfinite(depth_data) & (depth_data > 0)
    valid_pixels = np.sum(ok)
    total_pixels = depth_data.size
    zero_pixels = np.sum(depth_data == 0)
    inf_pixels = np.sum(np.isinf(depth_data))
    valid_depth = depth_data[ok]
    if len(valid_depth) > 0:
        depth_min, depth_max, depth_mean = np.min(valid_depth), np.max(valid_depth), np.mean(valid_depth)
    else:
        depth_min = depth_max = depth_mean = 0.0
    return {
        "status": "valid",
        "valid_pixels": int(valid_pixels),
        "total_pixels": int(total_pixels),
        "valid_ratio": float(valid_pixels / total_pixels),
        "zero_pixels": int(zero_pixels),
        "inf_pixels": int(inf_pixels),
        "depth_range": [float(depth_min), float(depth_max)],
        "depth_mean": float(depth_mean),
    }


# ------------------------------------------------------------------------------------------
# f3: the text files of the depth map and the point cloud  (gcd.py:1688, 1752-1753)
# ------------------------------------------------------------------------------------------
def savetxt_fixed6(values: np.ndarray, header: Optional[str] = None) -> bytes:
    """The bytes the reference's own call writes: ``np.savetxt(path, depth_data, delimiter=' ', fmt='%.6f')``
    (gcd.py:1688) and, with ``header='x y z r g b', comments=''``, the point-cloud file (gcd.py:1752-1753).
    numpy is the implementation here exactly as it is in the reference, so this row needs no restatement."""
    import io

    buf = io.BytesIO()
    if header is None:
        np.savetxt(buf, values, delimiter=" ", fmt="%.6f")
    else:
        np.savetxt(buf, values, fmt="%.6f", delimiter=" ", header=header, comments="")
    return buf.getvalue()


# ------------------------------------------------------------------------------------------
# f4: depth visualisation  (gcd.py:1691-1709) and RGB -> BGR (gcd.py:1671)
# ------------------------------------------------------------------------------------------
def jet_lut_bgr() -> np.ndarray:
    """The 256 x 3 BGR table behind cv2.applyColorMap(..., cv2.COLORMAP_JET) (gcd.py:1702)."""
    import cv2

    return cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(-1, 1), cv2.COLORMAP_JET).reshape(256, 3)


def depth_colormap(depth_data: np.ndarray) -> np.ndarray:
    """Restatement of the inline block gcd.py:1691-1709 (it is not a function in the reference,
    so it cannot be AST-extracted; PINNED ONLY BY READING): JET image of the depth normalised over
    its valid (finite, > 0) pixels, black when none is valid."""
    import cv2

    depth_valid = depth_data[np.isfinite(depth_data) & (depth_data > 0)]
    if len(depth_valid) > 0:
        depth_min = np.min(depth_valid)
        depth_max = np.max(depth_valid)
        depth_normalized = np.zeros_like(depth_data, dtype=np.uint8)
        mask = np.isfinite(depth_data) & (depth_data > 0)
        depth_normalized[mask] = ((depth_data[mask] - depth_min) / (depth_max - depth_min + 1e-6) * 255).astype(np.uint8)
        return cv2.applyColorMap(depth_normalized, cv2.COLORMAP_JET)
    return np.zeros(depth_data.shape + (3,), dtype=np.uint8)


def rgb_to_bgr(rgb_image: np.ndarray) -> np.ndarray:
    """gcd.py:1671: cv2.cvtColor(rgb_image[..., :3], cv2.COLOR_RGB2BGR)."""
    import cv2

    return cv2.cvtColor(np.ascontiguousarray(rgb_image[..., :3]), cv2.COLOR_RGB2BGR)
