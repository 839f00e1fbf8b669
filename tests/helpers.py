"""Shared test helpers: host tables from synthetic frames, the oracle pipeline, comparisons."""
from __future__ import annotations

import numpy as np

from constructionsceneposeestimation_b200 import classes
from oracle import labels as O

PX_ATOL = 1e-4   # north_star: float projections within 1e-4 px
REL_TOL = 1e-5   # north_star: poses within 1e-5 relative
# Euler angles against the REFERENCE's own bboxDict_to_transform (gcd.py:553-584), whose SVD runs in float32: the
# f64 polar factor of the kernel / oracle sits within 3.9e-6 degrees of it on the golden records, so the bar is
# 1e-5 degrees absolute + 1e-5 relative
EULER_REF_ATOL = 1e-5
# kernel against oracle (both f64, different algorithms for the polar factor): 1e-5 relative + 1e-9 degrees
EULER_ATOL = 1e-9


def host_tables(frames, split_people=True, fallback="first_mesh"):
    """(lut [B,L], obj_record [B,N], slot_class [B,N], records [B,R], cam [B,24], objects) as numpy,
    built with the PRODUCT's host logic (classes.py) and the ORACLE's camera packing.  ``fallback="union"``: the U
    object-level records of a frame sit behind the batch's own R records (records is [B, R+U], filled by the
    oracle's union_records; ``host_tables.union`` keeps (base, offsets, members) of the last call for kernel tests)."""
    res = classes.ObjectRootResolver(split_people=split_people)
    per = []
    for fr in frames:
        paths = fr["bounding_box_3d"]["info"]["primPaths"]
        objs = classes.aggregate_objects(paths, res)
        rec_idx = classes.record_index_for(objs, paths, fallback)
        mapping = classes.id_to_slot(fr["instance_segmentation"]["info"]["idToLabels"], objs, res)
        plan = classes.union_members(objs, paths) if fallback == "union" else []
        per.append((objs, rec_idx, mapping, plan))
    B = len(frames)
    N = max(1, max(len(p[0]) for p in per))
    R0 = max(1, max(len(fr["bounding_box_3d"]["data"]) for fr in frames))
    u_off, u_mem, U = classes.pack_union([p[3] for p in per])
    R = R0 + U
    L = max(1, max((max(p[2].keys()) if p[2] else 0) for p in per) + 1)
    lut = np.full((B, L), -1, dtype=np.int32)
    obj_record = np.full((B, N), -1, dtype=np.int32)
    slot_class = np.full((B, N), -1, dtype=np.int32)
    records = np.zeros((B, R), dtype=O.BBOX3D_DTYPE)
    cam = np.zeros((B, O.CAM_STRIDE))
    for i, (fr, (objs, rec_idx, mapping, plan)) in enumerate(zip(frames, per)):
        for k, v in mapping.items():
            lut[i, k] = v
        obj_record[i, : len(objs)] = rec_idx
        for u, (slot, _) in enumerate(plan):   # re-base: union records follow the batch's R0 own records
            obj_record[i, slot] = (R0 + u) | classes.RECORD_APPROX_BIT
        slot_class[i, : len(objs)] = [o.class_id for o in objs]
        r = fr["bounding_box_3d"]["data"]
        records[i, : len(r)] = r
        cam[i] = O.pack_camera(fr["camera_pose"], fr["camera_params"])
    host_tables.union = (R0, u_off, u_mem, records.copy())   # records BEFORE the union pass
    if U:
        records = O.union_records(records, R0, u_off, u_mem)
    return lut, obj_record, slot_class, records, cam, [p[0] for p in per]


def oracle_pipeline(frames, min_pixels=1, tol=0.15, frame_base=0, **kw):
    lut, obj_record, slot_class, records, cam, objects = host_tables(frames, **kw)
    mask = np.stack([fr["instance_segmentation"]["data"] for fr in frames])
    H, W = mask.shape[1:]
    N = obj_record.shape[1]
    scan = O.mask_scan(mask, lut, N)
    uv, z, pose, loose, flags = O.project_objects(records, obj_record, cam)
    recs, n_out, hist = O.emit(scan, uv, z, pose, loose, flags, slot_class, H, W, min_pixels, frame_base)
    out = dict(lut=lut, obj_record=obj_record, slot_class=slot_class, records=records, cam=cam, objects=objects,
               mask=mask, scan=scan, uv=uv, z=z, pose=pose, loose=loose, flags=flags, recs=recs, n_out=n_out,
               hist=hist)
    if all(fr.get("skeleton_data") is not None for fr in frames):
        joints = np.stack([fr["skeleton_data"]["globalTranslations"] for fr in frames])
        if joints.shape[1] > 0:
            depth = np.stack([fr["distance_to_image_plane"] for fr in frames])
            out["kp"], out["kz"], out["vis"] = O.keypoints(joints, depth, cam, tol)
            out["joints"], out["depth"] = joints, depth
    return out


def assert_pose_close(got: np.ndarray, want: np.ndarray, valid: np.ndarray):
    """pose blocks [.., 16]: translation/centre/size to 1e-5 relative, quaternion to 1e-5 of its unit norm, Euler
    angles to 1e-5 RELATIVE (+ 1e-9 degrees) — north_star's bar, not a fraction of the 180-degree range."""
    g, w = got[valid], want[valid]
    for sl in (slice(0, 3), slice(7, 10), slice(10, 13)):
        scale = np.maximum(np.linalg.norm(w[:, sl], axis=1, keepdims=True), 1e-12)
        assert np.all(np.abs(g[:, sl] - w[:, sl]) <= REL_TOL * scale + 1e-12), f"pose fields {sl}"
    assert np.all(np.abs(g[:, 3:7] - w[:, 3:7]) <= REL_TOL), "quaternion"
    de = np.abs(g[:, 13:16] - w[:, 13:16])
    de = np.minimum(de, 360.0 - de)  # +-180 wrap
    assert np.all(de <= EULER_ATOL + REL_TOL * np.abs(w[:, 13:16])), f"euler max diff {de.max()}"


def assert_records_equal(got: np.ndarray, want: np.ndarray, pose_tol=True):
    """Emitted records: every integer field bit-exact; float32 ratios to 1e-5 relative; pixel
    coordinates to 1e-4 px; poses as in assert_pose_close."""
    assert got.shape == want.shape
    for name in ("frame", "inst_idx", "class_id", "count", "x_min", "y_min", "x_max", "y_max", "flags", "loose"):
        assert np.array_equal(got[name], want[name]), name
    for name in ("occlusion", "fill", "truncation", "visible_frac", "yolo"):
        assert np.allclose(got[name], want[name], rtol=REL_TOL, atol=1e-7, equal_nan=True), name
    assert np.allclose(got["uv"], want["uv"], rtol=0, atol=PX_ATOL, equal_nan=True) or \
        np.allclose(got["uv"], want["uv"], rtol=REL_TOL, atol=PX_ATOL, equal_nan=True), "uv"
    assert np.allclose(got["z"], want["z"], rtol=REL_TOL, atol=1e-9, equal_nan=True), "z"
    if pose_tol and len(got):
        valid = (want["flags"] & O.OBJ_POSE_VALID) != 0
        assert_pose_close(got["pose"], want["pose"], valid)
