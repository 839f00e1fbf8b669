"""Pin the oracle (and the product's host logic) to the reference.

Two layers: (1) frozen fixtures under tests/golden/ that tests/golden/make_golden.py produced by
RUNNING the reference's own functions — these run everywhere, including the GPU box;
(2) where /root/reference exists (the build container), live comparison against the
AST-extracted reference functions on fresh seeded inputs.
"""
import contextlib
import io
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import labels as O
from tests import helpers
from oracle import reference_extract

GOLD = Path(__file__).resolve().parent / "golden"
needs_reference = pytest.mark.skipif(not reference_extract.available(), reason="/root/reference not on this machine")


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ------------------------------------------------------------------ R1 / R2: prim path -> object root, class
def test_object_roots_golden():
    from constructionsceneposeestimation_b200 import classes

    g = json.loads((GOLD / "object_roots.json").read_text())
    assert classes.CLASS_TABLE == g["construction_class"]
    assert list(classes.CLASS_TABLE) == list(g["construction_class"])  # key order drives the generic fallback
    assert {k: list(v) for k, v in classes.CRANE_CHILD_PARTS.items()} == g["crane_child_map"]
    res = classes.ObjectRootResolver()
    for path, want in zip(g["paths"], g["without_map"]):
        assert list(res.resolve(path)) == want, path
    res = classes.ObjectRootResolver({k: tuple(v) for k, v in g["crane_part_map"].items()})
    for path, want in zip(g["paths"], g["with_map"]):
        assert list(res.resolve(path)) == want, path


@needs_reference
def test_object_roots_live_fuzz():
    """Random recombinations of real path fragments: product resolver == reference get_object_root."""
    from constructionsceneposeestimation_b200 import classes

    ref = reference_extract.load()
    reference_extract.set_crane_part_map({})
    rng = np.random.default_rng(0)
    frags = ["World", "GroundPlane", "Tree", "Tree_03", "tree", "Cone001", "Cone001_02", "cone001", "DHGen", "dhgen_x",
             "SkelRoot", "tn__Pk7501SLD_PNR3879_fPM", "S104GG03A_SW", "s104kz02ka_sw", "tn__S104EKB_AS_SW_jJ7", "pk7",
             "boom", "mast", "Base", "teleskop", "tn__09684481_", "09684481", "Fencing_height_2,_07", "fencing_height_x",
             "Construction_Site_A_Fencing_height_12", "mesh", "Mesh_0", "human", "crane", "cranebase", "dumper", "fence",
             "construction_site", "trafficcone", "x", ""]
    res = classes.ObjectRootResolver()
    for _ in range(4000):
        n = int(rng.integers(1, 7))
        path = "/" + "/".join(frags[int(i)] for i in rng.integers(0, len(frags), n))
        if rng.uniform() < 0.1:
            path = path.lower()
        assert res.resolve(path) == tuple(ref.get_object_root(path)), path


@needs_reference
def test_aggregation_matches_reference_loop():
    """R2: grouping + inst_idx order == the loop at gcd.py:1858-1886 (restated inline from the
    reference's get_object_root, since that loop lives inside generate_data())."""
    from constructionsceneposeestimation_b200 import classes, synthetic

    ref = reference_extract.load()
    reference_extract.set_crane_part_map({})
    frame = synthetic.make_frame(synthetic.SceneSpec(320, 180, 40, 3, 17, config_id=5, split_people=False), 0)
    paths = frame["bounding_box_3d"]["info"]["primPaths"]
    roots = {}
    for pth in paths:  # gcd.py:1864-1875
        root, name, cid = ref.get_object_root(pth)
        if root is not None:
            roots.setdefault(root, {"class_id": cid, "class_name": name, "mesh_paths": []})["mesh_paths"].append(pth)
    objs = classes.aggregate_objects(paths, classes.ObjectRootResolver())
    assert [o.prim_path for o in objs] == list(roots)
    for i, (o, (root, info)) in enumerate(zip(objs, roots.items())):
        assert (o.inst_idx, o.class_id, o.class_name, o.mesh_paths) == (i, info["class_id"], info["class_name"],
                                                                       info["mesh_paths"])
    # record lookup: primPaths.index(root) first, crane parts fall back to their meshes (gcd.py:1934, 1953-1975)
    idx_ref = classes.record_index_for(objs, paths, "reference")
    idx_ext = classes.record_index_for(objs, paths, "first_mesh")
    for o, a, b in zip(objs, idx_ref, idx_ext):
        if o.actual_prim_path in paths and "#" not in o.prim_path:
            assert a == b == paths.index(o.actual_prim_path)
        elif "#" in o.prim_path:
            assert a == b and paths[a] in o.mesh_paths
        else:
            # a mesh record standing in for a multi-mesh object is marked as approximate (bit 30)
            assert a == -1 and paths[b & ~classes.RECORD_APPROX_BIT] == o.mesh_paths[0]
            assert bool(b & classes.RECORD_APPROX_BIT) == (len(o.mesh_paths) > 1)


def test_frame_golden_aggregation_and_record_lookup():
    """R2 frozen (runs without /root/reference): grouping, inst_idx order, mesh lists and the record lookup of
    gcd.py:1858-1886 / 1924-1975, as the reference's own functions produced them for one frame
    (tests/golden/make_golden.py::make_frame_golden)."""
    from constructionsceneposeestimation_b200 import classes
    g = json.loads((GOLD / "frame_label.json").read_text(encoding="utf-8"))
    paths = g["prim_paths"]
    objs = classes.aggregate_objects(paths, classes.ObjectRootResolver())
    assert len(objs) == len(g["object_list"])
    for o, want in zip(objs, g["object_list"]):
        assert (o.inst_idx, o.class_id, o.class_name, o.prim_path, o.mesh_count, o.mesh_paths) == \
            (want["inst_idx"], want["class_id"], want["class_name"], want["prim_path"], want["mesh_count"], want["mesh_paths"])
    assert classes.record_index_for(objs, paths, "reference") == g["record_index"]
    ext = classes.record_index_for(objs, paths, "first_mesh")
    for o, a, b in zip(objs, g["record_index"], ext):
        if a >= 0:
            assert b == a                                   # the reference's own rule wins wherever it finds a record
        else:
            assert paths[b & ~classes.RECORD_APPROX_BIT] == o.mesh_paths[0]
            assert bool(b & classes.RECORD_APPROX_BIT) == (o.mesh_count > 1)
    # the objects[] of the label file are the oracle's bbox_to_transform of those records
    recs = np.load(GOLD / "frame_records.npz")["records"]
    label = json.loads(g["label_text"])
    posed = [(o, i) for o, i in zip(objs, g["record_index"]) if i >= 0]
    assert label["num_objects"] == len(posed) == len(label["objects"])
    for (o, i), entry in zip(posed, label["objects"]):
        c, s, e = O.bbox_to_transform(recs[i])
        assert (entry["inst_idx"], entry["class_id"], entry["class_name"], entry["prim_path"]) == \
            (o.inst_idx, o.class_id, o.class_name, o.prim_path)
        assert np.allclose(c, entry["center"], rtol=1e-12, atol=1e-12) and np.allclose(s, entry["size"], rtol=1e-12, atol=1e-12)
        de = np.abs(np.asarray(e) - np.asarray(entry["rotation"]))
        assert np.all(np.minimum(de, 360 - de) <= helpers.EULER_REF_ATOL + helpers.REL_TOL * np.abs(entry["rotation"]))


# ------------------------------------------------------------------ R3: bbox record -> centre / size / euler
def test_bbox_to_transform_golden():
    g = np.load(GOLD / "bbox_to_transform.npz")
    for i, rec in enumerate(g["records"]):
        c, s, e = O.bbox_to_transform(rec)
        assert np.allclose(c, g["center"][i], rtol=1e-12, atol=1e-12)
        assert np.allclose(s, g["size"][i], rtol=1e-12, atol=1e-12)
        de = np.abs(np.asarray(e) - g["euler"][i])
        # f32 LAPACK may differ across numpy builds, but only in the last bits of a float32 rotation
        assert np.all(np.minimum(de, 360 - de) <= helpers.EULER_REF_ATOL + helpers.REL_TOL * np.abs(g["euler"][i]))
    meta = json.loads((GOLD / "META.json").read_text())
    if meta["numpy"] == np.__version__:  # same numpy build as the fixture: bit-identical
        for i, rec in enumerate(g["records"]):
            c, s, e = O.bbox_to_transform(rec)
            assert c == list(g["center"][i]) and s == list(g["size"][i]) and e == list(g["euler"][i])
    # mirrored transform: scipy raises (the reference's caller swallows it, gcd.py:1949)
    bad = g["records"][5].copy()
    bad["transform"][0, :3] *= -1
    with pytest.raises(ValueError):
        O.bbox_to_transform(bad)


@needs_reference
def test_bbox_to_transform_live():
    ref = reference_extract.load()
    from tests.golden.make_golden import random_records

    recs = random_records(np.random.default_rng(99), 40)
    for rec in recs:
        assert O.bbox_to_transform(rec) == ref.bboxDict_to_transform(rec)


def test_project_objects_pose_consistent_with_r3():
    """The [SPEC] pose block carries R3's centre/size/euler: check the f64 restatement inside
    project_objects against the reference restatement on the golden records."""
    g = np.load(GOLD / "bbox_to_transform.npz")
    recs = g["records"][None]
    obj_record = np.arange(recs.shape[1], dtype=np.int32)[None]
    cam = O.pack_camera([1, 2, 3, 0.1, 0.2, 0.3, 0.9], {"focal_length": 12.0, "horizontal_aperture": 25.0,
                                                        "vertical_aperture": 14.0625, "width": 1280, "height": 720})[None]
    _, _, pose, _, flags = O.project_objects(recs, obj_record, cam)
    assert np.all(flags[0] & O.OBJ_POSE_VALID)
    assert np.allclose(pose[0, :, 7:10], g["center"], rtol=1e-6, atol=1e-9)
    assert np.allclose(pose[0, :, 10:13], g["size"], rtol=1e-6, atol=1e-9)
    de = np.abs(pose[0, :, 13:16] - g["euler"])
    assert np.all(np.minimum(de, 360 - de) <= helpers.EULER_REF_ATOL + helpers.REL_TOL * np.abs(g["euler"]))
    q = pose[0, :, 3:7]
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-12) and np.all(q[:, 3] >= 0)


# ------------------------------------------------------------------ f1: depth -> point cloud
def test_pointcloud_golden():
    g = np.load(GOLD / "pointcloud.npz")
    params = json.loads(str(g["params"]))
    pose = list(g["pose"])
    for rgb, key, prm in ((g["rgb"], "out", params), (g["dark"], "out_dark", params), (g["rgb"], "out_defaults", {})):
        got = O.depth_to_pointcloud(g["depth"], rgb, prm, pose)
        assert got.shape == g[key].shape
        assert np.array_equal(got[:, 3:], g[key][:, 3:])
        assert np.allclose(got[:, :3], g[key][:, :3], rtol=1e-12, atol=1e-12)
    assert O.depth_to_pointcloud(np.full((4, 4), np.inf, dtype=np.float32), g["rgb"][:4, :4], params, pose) is None


@needs_reference
def test_pointcloud_live():
    ref = reference_extract.load()
    rng = np.random.default_rng(3)
    depth = rng.uniform(-1, 300, (30, 44)).astype(np.float32)
    depth[rng.uniform(size=depth.shape) < 0.3] = np.inf
    rgb = rng.integers(0, 256, (30, 44, 3), dtype=np.uint8)
    params = {"horizontal_aperture": 20.0, "vertical_aperture": 13.6, "focal_length": 15.0, "width": 44, "height": 30}
    pose = [1.0, 2.0, 3.0, 0.5, -0.5, 0.5, 0.5]
    with _quiet():
        want = ref.depth_to_pointcloud_with_rgb(depth, rgb, params, pose)
    assert np.array_equal(O.depth_to_pointcloud(depth, rgb, params, pose), want)


# ------------------------------------------------------------------ f2: depth statistics
def test_depth_stats_golden():
    g = np.load(GOLD / "depth_stats.npz")
    want = json.loads(str(g["results"]))
    for name, ref in want.items():
        got = O.depth_stats(g[name])
        for k in ("valid_pixels", "total_pixels", "zero_pixels", "inf_pixels", "depth_range"):
            assert got[k] == ref[k], (name, k)
        assert got["valid_ratio"] == ref["valid_ratio"]
        assert abs(got["depth_mean"] - ref["depth_mean"]) <= 1e-6 * max(1.0, abs(ref["depth_mean"]))


@needs_reference
def test_depth_stats_live(tmp_path):
    ref = reference_extract.load()
    rng = np.random.default_rng(8)
    d = rng.uniform(0, 100, (50, 70)).astype(np.float32)
    d[d < 10] = 0
    d[d > 90] = np.inf
    with _quiet():
        logger = ref.DataQualityLogger(str(tmp_path))
        logger.log_frame_start(0, [0, 0, 0])
        logger.log_depth(True, d)
    assert O.depth_stats(d) == logger.current_frame["depth"]


# ------------------------------------------------------------------ R5 / R4: camera
def test_intrinsics_follow_reference_formula():
    from constructionsceneposeestimation_b200 import camera

    # script set-up: f = 12, ha = 25 (gcd.py:1442-1443) -> fx = 614.4 @1280, 921.6 @1920, 1843.2 @3840
    for w, h, fx in ((1280, 720, 614.4), (1920, 1080, 921.6), (3840, 2160, 1843.2)):
        p = camera.camera_params(w, h)
        got = camera.intrinsics(p)
        assert got == O.intrinsics(p)
        assert abs(got[0] - fx) < 1e-9 and abs(got[1] - fx) < 1e-9 and got[2:] == (w / 2.0, h / 2.0)
    assert camera.intrinsics({}, 640, 480) == O.intrinsics({}, 640, 480)  # fallback constants gcd.py:2047-2053
    p = camera.camera_params(1280, 720)
    assert p["vertical_aperture"] == 25.0 * (720 / 1280)                  # gcd.py:2038


def test_pack_camera_matches_scipy_path():
    from constructionsceneposeestimation_b200 import camera

    rng = np.random.default_rng(4)
    for _ in range(50):
        q = rng.normal(size=4) * rng.uniform(0.5, 2.0)   # not normalised: scipy normalises (gcd.py:681)
        pose = list(rng.uniform(-30, 30, 3)) + list(q)
        p = camera.camera_params(1920, 1080)
        a, b = camera.pack_camera(pose, p), O.pack_camera(pose, p)
        assert np.allclose(a, b, rtol=0, atol=1e-15) and a.shape == (O.CAM_STRIDE,)
    m = np.eye(4)
    m[:3, :3] = camera.quat_xyzw_to_matrix([0.1, 0.2, 0.3, 0.9]).T   # USD row-vector matrix of that rotation
    m[3, :3] = (4, 5, 6)
    t, r = camera.pose_from_usd_matrix(m)
    pose7 = O.camera_pose_from_usd_matrix(m)
    assert np.allclose(t, pose7[:3]) and np.allclose(camera.quat_xyzw_to_matrix(pose7[3:]), r, atol=1e-12)


# ------------------------------------------------------------------ R7: label JSON
def test_label_json_schema_golden(tmp_path):
    from constructionsceneposeestimation_b200 import formats

    g = json.loads((GOLD / "label_schema.json").read_text())
    p = tmp_path / "label.json"
    formats.dump_label_json(g["label"], p)
    assert p.read_text(encoding="utf-8") == g["text"]   # byte-identical to the reference's save_label_json


# ------------------------------------------------------------------ S1: two numpy formulations + the C restatement
@pytest.mark.parametrize("seed", range(4))
def test_mask_scan_formulations_agree(seed):
    rng = np.random.default_rng(seed)
    H, W, n_ids, N = int(rng.integers(1, 60)), int(rng.integers(1, 90)), 14, 9
    mask = rng.integers(0, n_ids + 2, size=(2, H, W), dtype=np.uint32)
    mask[0, : H // 2] = 3
    lut = np.concatenate([[-1, -1], rng.integers(-1, N + 2, n_ids)]).astype(np.int32)
    a, b = O.mask_scan(mask, lut, N, fast=False), O.mask_scan(mask, lut, N, fast=True)
    assert np.array_equal(a, b)
    from oracle import c_oracle

    if c_oracle.available():
        assert np.array_equal(a, c_oracle.mask_scan(mask, lut, N))


def test_pointcloud_annotator_rows_golden(tmp_path):
    """formats.pointcloud_annotator_rows == what the reference's save_pointcloud_with_rgb writes (gcd.py:715-769) for
    every payload shape it special-cases: RGBA / RGB / missing / empty / short / long / flat / two-column colours, a
    single flat point, float colours, malformed and empty xyz.  The golden texts were written by the reference's own
    function (make_golden.py); here np.savetxt prints our matrix — the device formatter is held to np.savetxt's
    bytes by the GPU tests."""
    import importlib.util
    from constructionsceneposeestimation_b200 import formats
    spec = importlib.util.spec_from_file_location("make_golden", GOLD / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = json.loads((GOLD / "pointcloud_annotator.json").read_text())
    cases = mg.pointcloud_annotator_cases()
    assert sorted(want) == sorted(name for name, _ in cases) and sum(v is None for v in want.values()) == 3
    for name, payload in cases:
        rows = formats.pointcloud_annotator_rows(payload)
        if want[name] is None:
            assert rows is None, name
            continue
        assert rows is not None and rows.ndim == 2 and rows.shape[1] in (5, 6), name
        path = tmp_path / f"{name}.txt"
        np.savetxt(path, rows, fmt="%.6f", delimiter=" ", header="x y z r g b", comments="")     # gcd.py:768-769
        assert path.read_text() == want[name], name
        # the float64 copy the writer formats on the device prints the same text
        np.savetxt(path, rows.astype(np.float64), fmt="%.6f", delimiter=" ", header="x y z r g b", comments="")
        assert path.read_text() == want[name], name
    assert formats.pointcloud_annotator_rows(None) is None and formats.pointcloud_annotator_rows({}) is None


@needs_reference
def test_pointcloud_annotator_rows_live(tmp_path):
    """The same comparison against the reference function executed here, on random payloads."""
    from constructionsceneposeestimation_b200 import formats
    ref = reference_extract.load()
    rng = np.random.default_rng(23)
    import contextlib
    import io
    for trial in range(40):
        n = int(rng.integers(1, 50))
        payload = {"data": (rng.normal(size=(n, 3)) * 30).astype(np.float32)}
        kind = trial % 4
        if kind == 0:
            payload["pointRgb"] = rng.integers(0, 256, (n, 4), dtype=np.uint8)
        elif kind == 1:
            payload["pointRgb"] = rng.integers(0, 256, (int(rng.integers(1, 60)), 3), dtype=np.uint8)
        elif kind == 2:
            payload["pointRgb"] = rng.random((n, 4)).astype(np.float32)
        path = tmp_path / "ref.txt"
        if path.exists():
            path.unlink()
        with contextlib.redirect_stdout(io.StringIO()):
            ref.save_pointcloud_with_rgb(payload, str(path))
        rows = formats.pointcloud_annotator_rows(payload)
        if not path.exists():
            assert rows is None
            continue
        mine = tmp_path / "mine.txt"
        np.savetxt(mine, rows, fmt="%.6f", delimiter=" ", header="x y z r g b", comments="")
        assert mine.read_text() == path.read_text(), trial
