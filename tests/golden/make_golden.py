"""Freeze golden vectors by RUNNING the reference's own functions (build container only).

    python tests/golden/make_golden.py

Reads /root/reference/generate_construction_data.py through oracle.reference_extract (AST
extraction, nothing copied), feeds seeded inputs to the reference functions on the hot path and
writes their outputs next to this script.  The GPU box has no /root/reference; tests there (and
everywhere) replay these fixtures.  numpy/scipy versions used are recorded in the fixtures.
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import scipy

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))

from oracle import reference_extract  # noqa: E402

BBOX3D_DTYPE = np.dtype(
    [("semanticId", "<u4"), ("x_min", "<f4"), ("y_min", "<f4"), ("z_min", "<f4"), ("x_max", "<f4"),
     ("y_max", "<f4"), ("z_max", "<f4"), ("transform", "<f4", (4, 4)), ("occlusionRatio", "<f4")]
)

CRANE = "/World/GroundPlane/tn__Pk7501SLD_PNR3879_fPM"
FENCE = ("/World/GroundPlane/Construction_Site_Construction_Zeppelin_Rental_GmbH_Metal_Construction_Site_"
         "Fencing_height_2,_")


def golden_paths():
    """Real scene patterns (gcd.py:128-141), every crane child of gcd.py:110-121, keyword
    fallbacks, case games and paths that match nothing."""
    paths = [
        f"{FENCE}03", f"{FENCE}03/Mesh_0", f"{FENCE}25/Geom/panel/mesh", f"{FENCE.lower()}07/mesh",
        "/World/GroundPlane/construction_site_fencing_height_2,_09/Mesh",
        "/World/Tree/Tree", "/World/Tree/Tree/trunk", "/World/Tree/Tree_01", "/World/Tree/Tree_11/leaves/mesh_3",
        "/world/tree/tree_05/x", "/World/Tree", "/Other/World/Tree/Tree_02/a/b",
        "/World/GroundPlane/Cone001", "/World/GroundPlane/Cone001/Cone001", "/World/GroundPlane/Cone001_01/Cone001",
        "/World/GroundPlane/Cone001_02/Cone001/mesh", "/World/GroundPlane/cone001_07/x", "/World/Props/TrafficCone_3/mesh",
        CRANE, f"{CRANE}/S104GG03A_SW/mesh", f"{CRANE}/S104S01KB_SW", f"{CRANE}/S104HZ01KA_SW/a/b",
        f"{CRANE}/S104H01KB_SW/x", f"{CRANE}/S104HZ02KA_SW/x", f"{CRANE}/S104KZ01KA_SW/x",
        f"{CRANE}/tn__S104EKB_AS_SW_jJ7/part_0", f"{CRANE}/S104KZ02KA_SW/part", f"{CRANE}/tn__HHK320KA_SW_lG/m",
        f"{CRANE}/tn__HHK319_SW_oD/m", f"{CRANE}/UnknownChild/mesh", f"{CRANE}/Some/boom_segment/mesh",
        f"{CRANE}/x/chassis_plate", f"{CRANE}/x/Drehwerk", f"{CRANE}/x/teleskop_1", f"{CRANE}/x/Arm",
        f"{CRANE.lower()}/s104gg03a_sw/mesh", "/World/Other/pk7_copy/mast/mesh", "/World/Other/PK7/whatever",
        f"{CRANE}/exact/mapped/mesh", f"{CRANE}/exact/mapped/other",
        "/World/GroundPlane/tn__09684481_", "/World/GroundPlane/tn__09684481_/body/mesh", "/World/x/09684481/y",
        "/World/GroundPlane/DHGen", "/World/GroundPlane/DHGen/SkelRoot/body", "/World/GroundPlane/DHGen_03/SkelRoot/body",
        "/World/people/dhgen_female_01/mesh", "/World/Characters/worker/SkelRoot/mesh", "/World/Human_01/mesh",
        "/World/Props/Dumper_old/mesh", "/World/Props/crane_hook", "/World/Props/CraneBoom_spare", "/World/a/cranebase",
        "/World/Props/fence_post", "/World/construction_site_sign", "/World/GroundPlane/SomeUnlabelledProp/mesh",
        "/World/GroundPlane", "", "/", "BACKGROUND", "UNLABELLED", "/World/Tree/Trees_are_green/mesh",
    ]
    return paths


def make_paths(ref):
    paths = golden_paths()
    crane_map = {f"{CRANE}/exact/mapped/mesh": ("craneboom", 8), f"{CRANE}/S104GG03A_SW/mesh": ("cranecolumn", 7)}
    out = {"paths": paths, "crane_part_map": {k: list(v) for k, v in crane_map.items()}}
    reference_extract.set_crane_part_map({})
    out["without_map"] = [list(ref.get_object_root(p)) for p in paths]
    reference_extract.set_crane_part_map(crane_map)
    out["with_map"] = [list(ref.get_object_root(p)) for p in paths]
    reference_extract.set_crane_part_map({})
    out["construction_class"] = dict(ref.construction_class)
    out["crane_child_map"] = {k: list(v) for k, v in ref.CRANE_PART_CHILD_MAP.items()}
    (HERE / "object_roots.json").write_text(json.dumps(out, indent=1, ensure_ascii=False))
    return len(paths)


def random_records(rng, n):
    from scipy.spatial.transform import Rotation

    recs = np.zeros(n, dtype=BBOX3D_DTYPE)
    for i in range(n):
        lo = rng.uniform(-3, 0, 3)
        hi = lo + rng.uniform(0.05, 6, 3)
        rot = Rotation.random(random_state=int(rng.integers(1 << 31))).as_matrix()
        scale = rng.uniform(0.3, 2.5, 3) if i % 3 else np.full(3, rng.uniform(0.5, 2.0))
        shear = np.eye(3) + (rng.normal(0, 0.05, (3, 3)) if i % 5 == 0 else 0)
        m = np.eye(4)
        m[:3, :3] = (rot @ np.diag(scale) @ shear).T   # USD row-vector convention
        m[3, :3] = rng.uniform(-25, 25, 3)
        recs[i]["semanticId"] = i % 6
        recs[i]["x_min"], recs[i]["y_min"], recs[i]["z_min"] = lo
        recs[i]["x_max"], recs[i]["y_max"], recs[i]["z_max"] = hi
        recs[i]["transform"] = m.astype(np.float32)
        recs[i]["occlusionRatio"] = rng.uniform()
    # axis-aligned, pure yaw, and a near-gimbal pitch
    recs[0]["transform"] = np.eye(4, dtype=np.float32)
    yaw = Rotation.from_euler("z", 37.0, degrees=True).as_matrix()
    recs[1]["transform"][:3, :3] = yaw.T.astype(np.float32)
    gim = Rotation.from_euler("xyz", [20.0, 89.5, -40.0], degrees=True).as_matrix()
    recs[2]["transform"][:3, :3] = gim.T.astype(np.float32)
    return recs


def make_transforms(ref):
    rng = np.random.default_rng(20261018)
    recs = random_records(rng, 64)
    centers, sizes, eulers = [], [], []
    for r in recs:
        c, s, e = ref.bboxDict_to_transform(r)
        centers.append(c), sizes.append(s), eulers.append(e)
    np.savez(HERE / "bbox_to_transform.npz", records=recs, center=np.array(centers), size=np.array(sizes),
             euler=np.array(eulers))
    # the mirrored case raises inside scipy (gcd.py:1949 swallows it)
    bad = recs[5].copy()
    bad["transform"][0, :3] *= -1
    try:
        ref.bboxDict_to_transform(bad)
        raised = False
    except Exception:
        raised = True
    return len(recs), raised


def make_pointcloud(ref):
    rng = np.random.default_rng(7)
    H, W = 24, 32
    depth = rng.uniform(0.2, 300.0, (H, W)).astype(np.float32)
    depth[rng.uniform(size=(H, W)) < 0.2] = np.inf
    depth[0, 0], depth[1, 1], depth[2, 2], depth[3, 3] = 0.0, -2.0, np.nan, 250.0
    rgb = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    params = {"horizontal_aperture": 25.0, "vertical_aperture": 25.0 * H / W, "focal_length": 12.0, "width": W,
              "height": H}
    pose = [3.0, -4.0, 2.5, 0.1825742, 0.3651484, 0.5477226, 0.7302967]
    out = ref.depth_to_pointcloud_with_rgb(depth, rgb, params, pose)
    dark = (rgb > 200).astype(np.uint8)            # max <= 1 -> the x255 rule (gcd.py:693)
    out_dark = ref.depth_to_pointcloud_with_rgb(depth, dark, params, pose)
    out_defaults = ref.depth_to_pointcloud_with_rgb(depth, rgb, {}, pose)   # script default intrinsics
    none = ref.depth_to_pointcloud_with_rgb(np.full((H, W), np.inf, dtype=np.float32), rgb, params, pose)
    assert none is None
    np.savez(HERE / "pointcloud.npz", depth=depth, rgb=rgb, dark=dark, pose=np.array(pose), out=out, out_dark=out_dark,
             out_defaults=out_defaults, params=json.dumps(params))
    return out.shape


def pointcloud_annotator_cases():
    """Payloads of Replicator's pointcloud annotator as save_pointcloud_with_rgb meets them (gcd.py:715-769)."""
    rng = np.random.default_rng(17)
    xyz = (rng.normal(size=(37, 3)) * np.array([12.0, 12.0, 3.0])).astype(np.float32)
    xyz[0] = (0.0, -0.0, 1e-7)
    xyz[1] = (123456.789, -0.0000005, 2.5)
    rgba = rng.integers(0, 256, (37, 4), dtype=np.uint8)
    return [
        ("rgba", {"data": xyz, "pointRgb": rgba}),
        ("rgb3", {"data": xyz, "pointRgb": rgba[:, :3].copy()}),
        ("no_rgb", {"data": xyz}),
        ("rgb_none", {"data": xyz, "pointRgb": None}),
        ("rgb_empty", {"data": xyz, "pointRgb": np.zeros((0,), dtype=np.uint8)}),
        ("rgb_short", {"data": xyz, "pointRgb": rgba[:20]}),
        ("rgb_long", {"data": xyz[:11], "pointRgb": rgba}),
        ("rgb_flat_garbage", {"data": xyz, "pointRgb": rgba.reshape(-1)}),
        ("rgb_two_columns", {"data": xyz, "pointRgb": rgba[:, :2].copy()}),
        ("single_point_flat", {"data": xyz[5], "pointRgb": rgba[5]}),
        ("single_point_flat_rgb3", {"data": xyz[6], "pointRgb": rgba[6, :3].copy()}),
        ("float_colours", {"data": xyz.astype(np.float64), "pointRgb": rng.random((37, 4))}),
        ("xyz_flat_garbage", {"data": xyz.reshape(-1)[:7], "pointRgb": rgba}),
        ("xyz_empty", {"data": np.zeros((0, 3), dtype=np.float32), "pointRgb": rgba}),
        ("xyz_none", {"data": None, "pointRgb": rgba}),
    ]


def make_pointcloud_annotator(ref):
    """The text save_pointcloud_with_rgb writes for each payload (None = it writes no file)."""
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, payload in pointcloud_annotator_cases():
            path = Path(tmp) / f"{name}.txt"
            ref.save_pointcloud_with_rgb(payload, str(path))
            out[name] = path.read_text() if path.exists() else None
    (HERE / "pointcloud_annotator.json").write_text(json.dumps(out, indent=1))
    return sum(v is not None for v in out.values())


def make_depth_stats(ref):
    rng = np.random.default_rng(11)
    cases = {}
    d0 = rng.uniform(0.5, 250.0, (45, 80)).astype(np.float32)
    d0[rng.uniform(size=d0.shape) < 0.2] = np.inf
    d0[:3, :5] = 0.0
    d0[10, 10] = np.nan
    d0[11, 11] = -np.inf
    d0[12, 12] = -3.0
    cases["mixed"] = d0
    cases["all_zero"] = np.zeros((9, 14), dtype=np.float32)
    cases["all_inf"] = np.full((9, 14), np.inf, dtype=np.float32)
    cases["big"] = rng.uniform(0.5, 250.0, (360, 640)).astype(np.float32)
    results = {}
    with tempfile.TemporaryDirectory() as tmp:
        logger = ref.DataQualityLogger(tmp)
        for name, d in cases.items():
            logger.log_frame_start(0, [0, 0, 0])
            logger.log_depth(True, d)
            results[name] = logger.current_frame["depth"]
    np.savez(HERE / "depth_stats.npz", **cases, results=json.dumps(results))
    return list(results)


# events replayed through the reference's DataQualityLogger and, in the tests, through FrameQualityLog;
# "depth" names a case of depth_stats.npz (None = annotator returned nothing)
QUALITY_EVENTS = [
    {"frame": 0, "cam": [1.0, 2.0, 3.0], "cloud": [True, 5000, ""], "rgb": [True, ""], "depth": "mixed", "labels": 7, "ok": True},
    {"frame": 1, "cam": [0.0, 0.0, 9.5], "retry": 1, "cloud": [False, 0, "annotator返回None"], "ok": False},
    {"frame": 1, "cam": [0.0, 0.0, 9.5], "retry": 2, "cloud": [False, 40, "少于阈值 100"], "rgb": [False, "camera.get_rgba()返回None或空"],
     "depth": "all_zero", "labels": 0, "ok": True},
    {"frame": 2, "cam": [-4.0, 2.5, 3.0], "rgb": [True, ""], "depth": "all_inf", "labels": 3, "ok": True},
    {"frame": 3, "cam": [-4.0, 2.5, 3.0], "rgb": [True, ""], "depth": None, "labels": 12, "ok": True},
    {"frame": 4, "cam": [8.0, 8.0, 1.0], "cloud": [True, 123456, ""], "depth": "big", "labels": 1, "ok": True},
]


def make_quality_log(ref):
    cases = np.load(HERE / "depth_stats.npz")
    with tempfile.TemporaryDirectory() as tmp:
        logger = ref.DataQualityLogger(tmp)
        for ev in QUALITY_EVENTS:
            logger.log_frame_start(ev["frame"], np.asarray(ev["cam"]))
            if "retry" in ev:
                logger.log_retry(ev["retry"])
            if "cloud" in ev:
                logger.log_pointcloud(*ev["cloud"])
            if "rgb" in ev:
                logger.log_rgb(*ev["rgb"])
            if "depth" in ev:
                if ev["depth"] is None:
                    logger.log_depth(False, reason="annotator返回None或空")
                else:
                    logger.log_depth(True, cases[ev["depth"]])
            if "labels" in ev:
                logger.log_labels(ev["labels"])
            logger.log_frame_end(ev["ok"])
        report = logger.save_summary()
        summary = json.loads((Path(tmp) / "generation_summary.json").read_text(encoding="utf-8"))
    issue_lines = [ln.strip() for ln in report.split("常见问题:\n")[1].splitlines() if ln.strip()]
    (HERE / "quality_log.json").write_text(json.dumps({"events": QUALITY_EVENTS, "summary": summary,
                                                       "issue_lines": issue_lines, "report": report},
                                                      ensure_ascii=False, indent=1))
    return len(summary["frame_logs"])


def make_label_json(ref):
    label = {
        "frame_id": 7, "camera_pose": [1.0, 2.0, 3.0, 0.0, 0.0, 0.0, 1.0],
        "camera_params": {"horizontal_aperture": 25.0, "vertical_aperture": 14.0625, "focal_length": 12.0,
                          "width": 1280, "height": 720},
        "objects": [{"inst_idx": 0, "class_id": 2, "class_name": "fence", "center": [1.5, -2.25, 1.0],
                     "size": [3.5, 0.1, 2.0], "rotation": [0.0, -0.0, 37.0], "prim_path": f"{FENCE}03"}],
        "instance_mask_shape": [720, 1280], "num_objects": 1, "class_mapping": dict(ref.construction_class),
    }
    with tempfile.TemporaryDirectory() as tmp:
        p = Path(tmp) / "label.json"
        ref.save_label_json(label, str(p))
        text = p.read_text(encoding="utf-8")
    (HERE / "label_schema.json").write_text(json.dumps({"label": label, "text": text}, ensure_ascii=False, indent=1))
    return len(text)


def make_frame_golden(ref):
    """R2 + the per-object loop + the label file of ONE frame, produced by the reference's own functions.

    The loops at gcd.py:1858-1886 (aggregation, inst_idx order) and gcd.py:1924-1975 (record lookup: primPaths.index(root),
    crane parts through their mesh paths) live inside generate_data() and cannot be extracted, so they are driven from
    here exactly as written there, calling the reference's get_object_root / bboxDict_to_transform / save_label_json for
    every value.  The USD-stage fallback (gcd.py:1977-2023) does not exist offline: objects it would serve are absent
    from pose_list, as they are when the stage lookup fails."""
    from constructionsceneposeestimation_b200 import synthetic

    reference_extract.set_crane_part_map({})
    spec = synthetic.SceneSpec(64, 48, 40, 3, 17, config_id=31, split_people=False)
    frame = synthetic.make_frame(spec, 0)
    prim_paths = list(frame["bounding_box_3d"]["info"]["primPaths"])
    records = frame["bounding_box_3d"]["data"]
    # a camera far outside the site looking at it: every box is in front of it
    cam_pose = [0.0, -220.0, 6.0, 0.7071067811865476, 0.0, 0.0, 0.7071067811865476]
    object_roots = {}
    for prim_path in prim_paths:                                           # gcd.py:1864-1875
        object_root, class_name, class_id = ref.get_object_root(prim_path)
        if object_root is not None:
            if object_root not in object_roots:
                object_roots[object_root] = {"class_id": class_id, "class_name": class_name, "mesh_paths": []}
            object_roots[object_root]["mesh_paths"].append(prim_path)
    object_list = []
    inst_idx = 0
    for object_root, obj_info in object_roots.items():                     # gcd.py:1877-1886
        object_list.append({"inst_idx": inst_idx, "class_id": obj_info["class_id"], "class_name": obj_info["class_name"],
                            "prim_path": object_root, "mesh_count": len(obj_info["mesh_paths"]),
                            "mesh_paths": obj_info["mesh_paths"]})
        inst_idx += 1
    pose_list, record_index = [], []
    for obj in object_list:                                                # gcd.py:1924-1975
        prim_path = obj["prim_path"]
        actual_prim_path = prim_path.split("#")[0] if "#" in prim_path else prim_path
        idx = -1
        try:
            idx = prim_paths.index(actual_prim_path)
        except ValueError:
            if "#" in prim_path:
                for mesh_path in obj.get("mesh_paths", []):
                    try:
                        idx = prim_paths.index(mesh_path)
                        break
                    except ValueError:
                        continue
        record_index.append(idx)
        if idx < 0:
            continue
        center, size, euler = ref.bboxDict_to_transform(records[idx])
        pose_list.append({"inst_idx": obj["inst_idx"], "class_id": obj["class_id"], "class_name": obj["class_name"],
                          "center": center, "size": size, "rotation": euler, "prim_path": prim_path})
    cam_params = {"horizontal_aperture": 25.0, "vertical_aperture": 25.0 * (48 / 64), "focal_length": 12.0,
                  "width": 64, "height": 48}                               # gcd.py:2036-2045
    label_data = {"frame_id": 5, "camera_pose": cam_pose, "camera_params": cam_params, "objects": pose_list,
                  "instance_mask_shape": [48, 64], "num_objects": len(pose_list),
                  "class_mapping": ref.construction_class}                 # gcd.py:2056-2064
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "label_000005.json"
        ref.save_label_json(label_data, str(path))
        text = path.read_text(encoding="utf-8")
    np.savez(HERE / "frame_records.npz", records=records)
    (HERE / "frame_label.json").write_text(json.dumps({
        "prim_paths": prim_paths, "id_to_labels": frame["instance_segmentation"]["info"]["idToLabels"],
        "camera_pose": cam_pose, "camera_params": cam_params, "frame_id": 5,
        "object_list": object_list, "record_index": record_index, "label_text": text}, ensure_ascii=False, indent=1))
    return len(object_list), len(pose_list)


def main():
    ref = reference_extract.load()
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # the reference functions print progress text
        n_paths = make_paths(ref)
        n_recs, raised = make_transforms(ref)
        pc_shape = make_pointcloud(ref)
        n_pc_files = make_pointcloud_annotator(ref)
        stats = make_depth_stats(ref)
        n_text = make_label_json(ref)
        n_quality = make_quality_log(ref)
        n_objects, n_posed = make_frame_golden(ref)
    meta = {"numpy": np.__version__, "scipy": scipy.__version__, "paths": n_paths, "records": n_recs,
            "mirrored_transform_raises": raised, "pointcloud_shape": list(pc_shape), "pointcloud_annotator_files": n_pc_files, "depth_cases": stats,
            "label_text_bytes": n_text, "quality_frames": n_quality, "frame_objects": n_objects,
            "frame_objects_with_record": n_posed, "source": str(reference_extract.REFERENCE_SCRIPT)}
    (HERE / "META.json").write_text(json.dumps(meta, indent=1))
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
