"""Seeded synthetic annotator frames in the shapes the reference's capture loop receives.

SURVEY §8d.  One frame is the dict a Replicator writer gets: ``instance_segmentation``
(uint32 mask + ``idToLabels``), ``distance_to_image_plane`` (float32, ``inf`` = sky,
gcd.py:318-321), ``bounding_box_3d`` (structured records + ``primPaths``,
gcd.py:1788-1790), ``camera_params`` / ``camera_pose`` in the reference's own field names
(gcd.py:2039-2045, 1599) and ``skeleton_data``.  Prim paths use the real scene's patterns
(gcd.py:128-141); the class mix mirrors world2.usd (24 fence panels, 12 trees, 3 cones, a
crane in parts, a dumper, people) scaled to the requested instance count.

Masks, boxes and depth agree with each other: every object's blob is painted inside the
projection of its 3D box, back to front, at the depth of the box centre.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from ._lib import BBOX3D_DTYPE
from .camera import camera_params as make_camera_params
from .camera import intrinsics, quat_xyzw_to_matrix

FENCE_PREFIX = ("/World/GroundPlane/Construction_Site_Construction_Zeppelin_Rental_GmbH_Metal_"
                "Construction_Site_Fencing_height_2,_")
CRANE_ROOT = "/World/GroundPlane/tn__Pk7501SLD_PNR3879_fPM"
DUMPER_ROOT = "/World/GroundPlane/tn__09684481_"
CRANE_CHILDREN = ("S104GG03A_SW", "S104HZ01KA_SW", "tn__S104EKB_AS_SW_jJ7", "S104KZ02KA_SW")

# rough half-extents (m) of the local boxes per class: x, y, z
_EXTENTS = {
    "cone": (0.2, 0.2, 0.35), "tree": (1.6, 1.6, 3.0), "fence": (1.75, 0.05, 1.0), "crane": (1.2, 1.0, 1.5),
    "dumper": (2.5, 1.2, 1.4), "human": (0.3, 0.25, 0.9),
}

# COCO-17 rest pose of a 1.75 m standing figure (x right, y forward, z up), metres
_COCO17 = np.array(
    [[0.00, 0.08, 1.65], [0.03, 0.10, 1.69], [-0.03, 0.10, 1.69], [0.08, 0.03, 1.66], [-0.08, 0.03, 1.66],
     [0.20, 0.0, 1.45], [-0.20, 0.0, 1.45], [0.28, 0.0, 1.15], [-0.28, 0.0, 1.15], [0.30, 0.05, 0.88],
     [-0.30, 0.05, 0.88], [0.12, 0.0, 0.95], [-0.12, 0.0, 0.95], [0.13, 0.02, 0.50], [-0.13, 0.02, 0.50],
     [0.13, 0.0, 0.08], [-0.13, 0.0, 0.08]], dtype=np.float64)


@dataclass
class SceneSpec:
    width: int = 1280
    height: int = 720
    num_instances: int = 20      # aggregated objects (label slots) aimed for
    num_people: int = 4
    num_joints: int = 17         # 17 = COCO subset, anything larger = interpolated full rig
    config_id: int = 1           # seeds: default_rng(1000 * config_id + frame)
    sparse_ids: bool = False     # instance ids spread up to 2**20 instead of dense 2..
    split_people: bool = True    # one object per person (False = the reference's single DHGen root)
    with_rgb: bool = False
    max_meshes: int = 3          # meshes (= instance ids) per object
    # camera ring around the site centre: (r_min, r_max, z_min, z_max) in metres, scaled with the site.
    # The default keeps the whole site in view from outside; SURVEY §8d's ring (4-12 m out, 1.6-3 m up,
    # gcd.py:790,857) puts the camera among the objects, which then fill the frame ("dense" configs)
    camera_ring: Tuple[float, float, float, float] = (30.0, 42.0, 7.0, 12.0)
    # see-through textures: fence panels become a wire mesh (one object pixel line every 4 px in x and y)
    # and tree crowns foliage (a hashed 60 % pixel pattern) — whatever lies behind shows through at pixel
    # scale, the most fragmented masks the reference's scene produces (24 fence panels, 12 trees,
    # gcd.py:128-141)
    textured: bool = False


def joint_template(num_joints: int) -> np.ndarray:
    """[J,3] rest pose: the COCO-17 points, then deterministic interpolations between them."""
    if num_joints <= 17:
        return _COCO17[:num_joints].copy()
    extra = []
    i = 0
    while len(extra) < num_joints - 17:
        a, b = _COCO17[i % 17], _COCO17[(i * 7 + 3) % 17]
        w = ((i * 37) % 89 + 5) / 100.0
        extra.append(a * (1.0 - w) + b * w)
        i += 1
    return np.vstack([_COCO17, np.array(extra)])


def _look_at_pose(pos: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Camera 7-vector in USD axes: -Z forward, +Y up (world up = +Z), xyzw quaternion."""
    fwd = target - pos
    fwd = fwd / np.linalg.norm(fwd)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right = right / np.linalg.norm(right)
    up = np.cross(right, fwd)
    r = np.stack([right, up, -fwd], axis=1)  # columns = camera x, y, z axes in world
    # matrix -> quaternion (trace branch order is irrelevant here, only validity)
    t = np.trace(r)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(r[2, 1] - r[1, 2]) / s, (r[0, 2] - r[2, 0]) / s, (r[1, 0] - r[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(r)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + r[i, i] - r[j, j] - r[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (r[j, i] + r[i, j]) / s
        q[k] = (r[k, i] + r[i, k]) / s
        q[3] = (r[k, j] - r[j, k]) / s
    return np.concatenate([pos, q / np.linalg.norm(q)])


def _object_catalogue(spec: SceneSpec) -> List[Tuple[str, str, List[str]]]:
    """[(kind, root_path_hint, mesh paths)] in scene order; number of ROOTS == num_instances."""
    n = spec.num_instances
    people = min(spec.num_people, n) if spec.split_people else (1 if spec.num_people and n else 0)
    rest = n - people
    n_dumper = 1 if rest >= 4 else 0
    n_crane = min(4, max(0, rest - n_dumper - 2)) if rest >= 8 else 0
    rest2 = rest - n_dumper - n_crane
    n_cone = max(0, rest2 // 8) if rest2 >= 3 else 0
    n_tree = rest2 * 3 // 10
    n_fence = rest2 - n_cone - n_tree
    cat: List[Tuple[str, str, List[str]]] = []
    for i in range(n_fence):
        root = f"{FENCE_PREFIX}{i + 3:02d}"
        cat.append(("fence", root, [f"{root}/Mesh_{m}" for m in range(1 + i % spec.max_meshes)]))
    for i in range(n_tree):
        root = "/World/Tree/Tree" if i == 0 else f"/World/Tree/Tree_{i:02d}"
        cat.append(("tree", root, [f"{root}/trunk", f"{root}/leaves"][: 1 + i % 2]))
    for i in range(n_cone):
        root = "/World/GroundPlane/Cone001" if i == 0 else f"/World/GroundPlane/Cone001_{i:02d}"
        cat.append(("cone", root, [f"{root}/Cone001"]))
    for i in range(n_crane):
        child = CRANE_CHILDREN[i]
        cat.append(("crane", f"{CRANE_ROOT}/{child}", [f"{CRANE_ROOT}/{child}/part_{m}" for m in range(1 + i % 2)]))
    if n_dumper:
        cat.append(("dumper", DUMPER_ROOT, [f"{DUMPER_ROOT}/body", f"{DUMPER_ROOT}/bed"]))
    if spec.split_people:
        for i in range(people):
            root = f"/World/GroundPlane/DHGen_{i:02d}"
            cat.append(("human", root, [f"{root}/SkelRoot/body"]))
    elif people:
        root = "/World/GroundPlane/DHGen"
        cat.append(("human", root, [f"{root}/SkelRoot/body_{i}" for i in range(max(1, spec.num_people))]))
    return cat


def _rot_z(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _rot_x(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def make_frame(spec: SceneSpec, frame: int = 0) -> Dict[str, object]:
    """One seeded annotator dict (see module docstring)."""
    rng = np.random.default_rng(1000 * spec.config_id + frame)
    W, H = spec.width, spec.height
    params = make_camera_params(W, H)
    fx, fy, cx, cy = intrinsics(params)

    # camera on a ring around the site, looking at its centre at roughly eye height
    ang = rng.uniform(0.0, 2.0 * np.pi)
    site = max(1.0, np.sqrt(spec.num_instances / 100.0))  # the site grows with the instance count
    r_lo, r_hi, z_lo, z_hi = spec.camera_ring
    rad = rng.uniform(r_lo, r_hi) * site
    pos = np.array([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(z_lo, z_hi) * site])
    target = np.array([rng.uniform(-3.0, 3.0), rng.uniform(-3.0, 3.0), rng.uniform(0.8, 2.0)])
    pose7 = _look_at_pose(pos, target)
    rcw = quat_xyzw_to_matrix(pose7[3:])

    cat = _object_catalogue(spec)
    n_obj = len(cat)

    # ---- bbox3d records: one per mesh path; some roots also get their own record ----------
    prim_paths: List[str] = []
    rec_rows = []
    obj_geom = []  # (centre_world, half extents scaled, rotation) for painting
    sem_id = {"cone": 0, "tree": 1, "fence": 2, "crane": 3, "dumper": 4, "human": 5}
    people_roots = []
    for oi, (kind, root, meshes) in enumerate(cat):
        he = np.array(_EXTENTS[kind]) * rng.uniform(0.8, 1.25, size=3)
        scale = rng.uniform(0.5, 1.6)
        yaw = rng.uniform(-np.pi, np.pi)
        tilt = rng.uniform(-0.06, 0.06) if kind != "human" else 0.0
        rot = _rot_z(yaw) @ _rot_x(tilt)
        base = np.array([rng.uniform(-20.0, 20.0) * site, rng.uniform(-20.0, 20.0) * site, 0.0])
        centre = base + np.array([0.0, 0.0, he[2] * scale])
        lo = (-he + rng.uniform(-0.05, 0.05, size=3)).astype(np.float32)
        hi = (he + rng.uniform(-0.05, 0.05, size=3)).astype(np.float32)
        m = np.eye(4)
        m[:3, :3] = (rot * scale).T  # USD row-vector convention: rows are the images of the local axes
        m[3, :3] = centre
        row = np.zeros((), dtype=BBOX3D_DTYPE)
        row["semanticId"] = sem_id[kind]
        row["x_min"], row["y_min"], row["z_min"] = lo
        row["x_max"], row["y_max"], row["z_max"] = hi
        row["transform"] = m.astype(np.float32)
        row["occlusionRatio"] = rng.uniform(0.0, 1.0)
        own_record = kind in ("tree", "fence", "dumper") and (oi % 3 != 1)
        if own_record:  # the root prim itself carries a record (found by primPaths.index(root), gcd.py:1934)
            prim_paths.append(root)
            rec_rows.append(row)
        for mp in meshes:
            prim_paths.append(mp)
            rec_rows.append(row)
        obj_geom.append((centre, he * scale, rot))
        if kind == "human":
            people_roots.append((base, yaw))
    records = np.array(rec_rows, dtype=BBOX3D_DTYPE) if rec_rows else np.zeros((0,), dtype=BBOX3D_DTYPE)

    # ---- instance ids: one per mesh path (+ one unmatched distractor) ------------------------
    mesh_list = [(oi, mp) for oi, (_, _, meshes) in enumerate(cat) for mp in meshes]
    n_ids = len(mesh_list) + 1
    if spec.sparse_ids:
        ids = 2 + np.sort(rng.choice((1 << 20) - 2, size=n_ids, replace=False))
        ids = ids[rng.permutation(n_ids)]
    else:
        ids = 2 + rng.permutation(n_ids)
    id_to_labels = {"0": "BACKGROUND", "1": "UNLABELLED"}
    mesh_ids: Dict[int, List[int]] = {}
    for (oi, mp), iid in zip(mesh_list, ids[:-1]):
        id_to_labels[str(int(iid))] = mp
        mesh_ids.setdefault(oi, []).append(int(iid))
    distractor_id = int(ids[-1])
    id_to_labels[str(distractor_id)] = "/World/GroundPlane/SomeUnlabelledProp/mesh"

    # ---- depth: ground plane z = 0 seen from the camera, sky = inf ---------------------------
    us = (np.arange(W, dtype=np.float32) + 0.5 - np.float32(cx)) / np.float32(fx)
    vs = -(np.arange(H, dtype=np.float32) + 0.5 - np.float32(cy)) / np.float32(fy)
    r32 = rcw.astype(np.float32)
    dir_z = r32[2, 0] * us[None, :] + r32[2, 1] * vs[:, None] - r32[2, 2]  # world z of ray (x, y, -1)
    with np.errstate(divide="ignore", invalid="ignore"):
        depth = np.where(dir_z < 0, np.float32(-pos[2]) / dir_z, np.float32(np.inf)).astype(np.float32)
    depth[depth > 250.0] = np.inf
    mask = np.zeros((H, W), dtype=np.uint32)

    # a flat distractor patch on the ground (its id maps to no object)
    px0, py0 = int(rng.integers(0, max(1, W - 40))), int(rng.integers(H // 2, max(H // 2 + 1, H - 20)))
    mask[py0:py0 + 17, px0:px0 + 37] = distractor_id

    # ---- paint objects back to front ---------------------------------------------------------
    corners = np.array([[(k & 1), (k >> 1) & 1, (k >> 2) & 1] for k in range(8)], dtype=np.float64) * 2.0 - 1.0
    order = []
    for oi, (centre, he, rot) in enumerate(obj_geom):
        pc = rcw.T @ (centre - pos)
        order.append((-pc[2], oi))
    for zc, oi in sorted(order, reverse=True):
        if zc <= 0.6:
            continue
        centre, he, rot = obj_geom[oi]
        pw = centre[None, :] + (corners * he[None, :]) @ rot.T
        pcs = (pw - pos[None, :]) @ rcw
        zz = -pcs[:, 2]
        if (zz <= 0.5).any():
            continue
        uu = cx + fx * pcs[:, 0] / zz
        vv = cy - fy * pcs[:, 1] / zz
        x0, x1 = int(np.floor(uu.min())), int(np.ceil(uu.max()))
        y0, y1 = int(np.floor(vv.min())), int(np.ceil(vv.max()))
        bx0, bx1, by0, by1 = max(x0, 0), min(x1, W), max(y0, 0), min(y1, H)
        if bx1 - bx0 < 1 or by1 - by0 < 1:
            continue
        ys = np.arange(by0, by1)[:, None]
        xs = np.arange(bx0, bx1)[None, :]
        kind = cat[oi][0]
        if kind in ("tree", "cone", "human", "crane"):  # ellipse inscribed in the projected box
            ex, ey = (x0 + x1) * 0.5, (y0 + y1) * 0.5
            rx, ry = max((x1 - x0) * 0.5, 0.5), max((y1 - y0) * 0.5, 0.5)
            blob = ((xs - ex) / rx) ** 2 + ((ys - ey) / ry) ** 2 <= 1.0
        else:  # inset rectangle
            ix, iy = (x1 - x0) // 10, (y1 - y0) // 10
            blob = (xs >= x0 + ix) & (xs < x1 - ix) & (ys >= y0 + iy) & (ys < y1 - iy)
        if spec.textured and kind == "fence":      # wire mesh
            blob = blob & (((xs & 3) == 0) | ((ys & 3) == 0))
        elif spec.textured and kind == "tree":     # foliage
            blob = blob & ((((xs * 73856093) ^ (ys * 19349663) ^ (oi * 83492791)) % 5) < 3)
        if not blob.any():
            continue
        mids = mesh_ids[oi]
        # meshes split the blob into horizontal bands
        band = np.minimum(((ys - y0) * len(mids)) // max(1, (y1 - y0)), len(mids) - 1)
        idimg = np.asarray(mids, dtype=np.uint32)[np.broadcast_to(band, blob.shape)]
        sub_m = mask[by0:by1, bx0:bx1]
        sub_d = depth[by0:by1, bx0:bx1]
        sub_m[blob] = idimg[blob]
        sub_d[blob] = np.float32(zc)

    # ---- skeletons -----------------------------------------------------------------------------
    P, J = spec.num_people, spec.num_joints
    joints = np.zeros((P, J, 3), dtype=np.float32)
    tmpl = joint_template(J)
    for p in range(P):
        if p < len(people_roots):
            base, yaw = people_roots[p]
        else:
            base, yaw = np.array([rng.uniform(-20, 20), rng.uniform(-20, 20), 0.0]), rng.uniform(-np.pi, np.pi)
        jitter = rng.normal(0.0, 0.01, size=tmpl.shape)
        joints[p] = ((tmpl + jitter) @ _rot_z(yaw).T + base).astype(np.float32)

    frame_dict: Dict[str, object] = {
        "frame_id": frame,
        "instance_segmentation": {"data": mask, "info": {"idToLabels": id_to_labels}},
        "distance_to_image_plane": depth,
        "bounding_box_3d": {"data": records, "info": {"primPaths": prim_paths}},
        "camera_params": params,
        "camera_pose": [float(v) for v in pose7],
        "skeleton_data": {"globalTranslations": joints},
    }
    if spec.with_rgb:
        rgb = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
        rgb[..., 3] = 255
        frame_dict["rgb"] = rgb
    return frame_dict


def make_batch(spec: SceneSpec, num_frames: int, first_frame: int = 0) -> List[Dict[str, object]]:
    return [make_frame(spec, first_frame + i) for i in range(num_frames)]


# BASELINE.json configs (C1..C4; C5 streams C2-shaped frames)
CONFIGS = {
    "c1": SceneSpec(1280, 720, 20, 4, 17, config_id=1),
    "c2": SceneSpec(1920, 1080, 100, 4, 17, config_id=2),
    "c3": SceneSpec(1920, 1080, 60, 50, 17, config_id=3),
    "c4": SceneSpec(3840, 2160, 500, 8, 17, config_id=4),
    # stress variants of c2 for the mask scan (not BASELINE configs): the camera on SURVEY §8d's ring, inside
    # the site, and the same with see-through fences / trees
    "c2_dense": SceneSpec(1920, 1080, 100, 4, 17, config_id=12, camera_ring=(4.0, 12.0, 1.6, 3.0)),
    "c2_textured": SceneSpec(1920, 1080, 100, 4, 17, config_id=13, camera_ring=(4.0, 12.0, 1.6, 3.0), textured=True),
}
