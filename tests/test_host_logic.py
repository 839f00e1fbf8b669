"""Host-side logic without a GPU: synthetic annotators, tables, formats, sharding (gloo, world 2)."""
import json
import os
import socket
from pathlib import Path

import numpy as np
import pytest

from constructionsceneposeestimation_b200 import classes, formats, sharding, synthetic
from oracle import labels as O
from tests import helpers

GOLDEN = Path(__file__).resolve().parent / "golden"


# ------------------------------------------------------------------ synthetic annotator frames
def test_synthetic_frame_shape_and_determinism():
    spec = synthetic.SceneSpec(320, 180, 20, 3, 17, config_id=1)
    a, b = synthetic.make_frame(spec, 5), synthetic.make_frame(spec, 5)
    mask, depth = a["instance_segmentation"]["data"], a["distance_to_image_plane"]
    assert mask.shape == (180, 320) and mask.dtype == np.uint32
    assert depth.shape == (180, 320) and depth.dtype == np.float32
    assert np.array_equal(mask, b["instance_segmentation"]["data"]) and a["camera_pose"] == b["camera_pose"]
    assert not np.array_equal(mask, synthetic.make_frame(spec, 6)["instance_segmentation"]["data"])
    recs, paths = a["bounding_box_3d"]["data"], a["bounding_box_3d"]["info"]["primPaths"]
    assert recs.dtype.itemsize == 96 and len(recs) == len(paths)
    labels = a["instance_segmentation"]["info"]["idToLabels"]
    assert labels["0"] == "BACKGROUND" and labels["1"] == "UNLABELLED"
    assert set(np.unique(mask)) <= {int(k) for k in labels}            # every painted id is labelled
    assert np.isinf(depth).any() and np.isfinite(depth).any()           # sky + ground
    assert a["skeleton_data"]["globalTranslations"].shape == (3, 17, 3)
    assert len(classes.aggregate_objects(paths, classes.ObjectRootResolver(split_people=True))) == 20


def test_synthetic_masks_agree_with_boxes():
    """Every object's tight 2D box lies inside the projection of its 3D box (what the painter promises)."""
    frames = synthetic.make_batch(synthetic.SceneSpec(480, 270, 30, 3, 17, config_id=2), 2)
    o = helpers.oracle_pipeline(frames)
    seen = 0
    for f in range(2):
        for r in o["recs"][f, : o["n_out"][f]]:
            if not (r["flags"] & O.OBJ_ALL_FRONT):
                continue
            u, v = r["uv"][:, 0], r["uv"][:, 1]
            assert r["x_min"] >= np.floor(u.min()) - 1 and r["x_max"] <= np.ceil(u.max()) + 1
            assert r["y_min"] >= np.floor(v.min()) - 1 and r["y_max"] <= np.ceil(v.max()) + 1
            assert 0.0 <= r["occlusion"] <= 1.0 and 0.0 < r["fill"] <= 1.0
            seen += 1
    assert seen >= 10


def test_sparse_ids_and_reference_people_rule():
    spec = synthetic.SceneSpec(320, 180, 16, 4, 17, config_id=3, sparse_ids=True, split_people=False)
    fr = synthetic.make_frame(spec, 0)
    ids = [int(k) for k in fr["instance_segmentation"]["info"]["idToLabels"]]
    assert max(ids) > 10_000                                             # spread up to 2**20
    objs = classes.aggregate_objects(fr["bounding_box_3d"]["info"]["primPaths"], classes.ObjectRootResolver())
    humans = [o for o in objs if o.class_name == "human"]
    assert len(humans) == 1 and humans[0].prim_path == "/World/GroundPlane/DHGen" and humans[0].mesh_count == 4


# ------------------------------------------------------------------ tables
def test_id_to_slot_and_record_lookup():
    res = classes.ObjectRootResolver()
    crane = classes.CRANE_ROOT
    paths = ["/World/Tree/Tree_01", "/World/Tree/Tree_01/trunk", "/World/GroundPlane/Cone001_02/Cone001",
             f"{crane}/S104GG03A_SW/m0", f"{crane}/S104S01KB_SW/m1", "/World/Unknown/thing"]
    objs = classes.aggregate_objects(paths, res)
    assert [(o.prim_path, o.class_id, o.mesh_count) for o in objs] == [
        ("/World/Tree/Tree_01", 1, 2), ("/World/GroundPlane/Cone001_02", 0, 1), (f"{crane}#cranebase", 6, 2)]
    assert classes.record_index_for(objs, paths, "reference") == [0, -1, 3]
    assert classes.record_index_for(objs, paths, "first_mesh") == [0, 2, 3]
    with pytest.raises(ValueError):
        classes.record_index_for(objs, paths, "nope")
    labels = {"0": "BACKGROUND", "1": "UNLABELLED", "7": "/World/Tree/Tree_01/leaves", 9: paths[3],
              "11": {"class": "cone"}, "12": {"primPath": paths[2]}, "13": "/World/Tree/Tree_77/x"}
    assert classes.id_to_slot(labels, objs, res) == {7: 0, 9: 2, 12: 1}


# ------------------------------------------------------------------ formats
def _oracle_records():
    frames = synthetic.make_batch(synthetic.SceneSpec(320, 180, 14, 2, 17, config_id=4), 1)
    o = helpers.oracle_pipeline(frames)
    return frames[0], o, o["recs"][0, : o["n_out"][0]]


def test_reference_label_schema_and_extras(tmp_path):
    fr, o, recs = _oracle_records()
    kps = {int(recs[0]["inst_idx"]): formats.coco_keypoint_block(o["kp"][0, 0], o["vis"][0, 0])}
    lab = formats.reference_label(3, fr["camera_pose"], fr["camera_params"], 180, 320, recs, o["objects"][0], kps)
    assert list(lab) == ["frame_id", "camera_pose", "camera_params", "objects", "instance_mask_shape", "num_objects",
                         "class_mapping"]                                     # gcd.py:2056-2064
    assert lab["instance_mask_shape"] == [180, 320] and lab["num_objects"] == len(recs)
    assert lab["class_mapping"] == classes.CLASS_TABLE
    obj = lab["objects"][0]
    assert list(obj)[:7] == ["inst_idx", "class_id", "class_name", "center", "size", "rotation", "prim_path"]
    assert obj["bbox_2d_tight"] == [int(recs[0][k]) for k in ("x_min", "y_min", "x_max", "y_max")]
    assert len(obj["bbox_3d_projected"]) == 8 and "keypoints" in obj
    p = tmp_path / "l.json"
    formats.dump_label_json(lab, p)
    assert json.loads(p.read_text(encoding="utf-8"))["objects"][0]["prim_path"] == obj["prim_path"]


def test_yolo_and_coco():
    fr, o, recs = _oracle_records()
    lines = formats.yolo_lines(recs)
    assert len(lines) == len(recs)
    c, cx, cy, w, h = lines[0].split()
    assert int(c) == recs[0]["class_id"] and 0 < float(w) <= 1 and 0 < float(h) <= 1
    assert abs(float(cx) - (recs[0]["x_min"] + recs[0]["x_max"] + 1) / 2 / 320) < 1e-5
    anns = formats.coco_annotations(recs, image_id=3, first_ann_id=10)
    assert [a["id"] for a in anns] == list(range(10, 10 + len(recs)))
    a0 = anns[0]
    assert a0["bbox"] == [int(recs[0]["x_min"]), int(recs[0]["y_min"]), int(recs[0]["x_max"] - recs[0]["x_min"] + 1),
                          int(recs[0]["y_max"] - recs[0]["y_min"] + 1)]
    assert a0["area"] == recs[0]["count"] and a0["iscrowd"] == 0
    blk = formats.coco_keypoint_block(o["kp"][0, 0], o["vis"][0, 0])
    assert len(blk["keypoints"]) == 17 * 3 and blk["num_keypoints"] == int((o["vis"][0, 0] > 0).sum())
    assert [c["name"] for c in formats.coco_categories()][:6] == ["trafficcone", "tree", "fence", "crane", "dumper", "human"]


# ------------------------------------------------------------------ sharding
@pytest.mark.parametrize("frames,world", [(100_000, 8), (64, 3), (5, 8), (0, 4), (257, 2)])
def test_frame_ranges_partition(frames, world):
    ranges = [sharding.frame_range(r, world, frames) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == frames
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))              # contiguous, no gap, no overlap
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frame_range(world, world, frames)


def test_rig_ranges_and_batches():
    pairs = [p for r in range(8) for p in sharding.rig_frame_range(r, 8, 10, 4)]
    assert pairs == [(i // 4, i % 4) for i in range(40)]                      # config 4: 4-camera rig flattened
    assert sharding.batches(3, 200, 64) == [(3, 67), (67, 131), (131, 195), (195, 200)]
    assert sharding.batches(5, 5, 64) == []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_worker(rank, world, port, n_frames, out_dir):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.frame_range(rank, world, n_frames)
    spec = synthetic.SceneSpec(256, 144, 12, 2, 17, config_id=7)
    frames = [synthetic.make_frame(spec, i) for i in range(lo, hi)]
    o = helpers.oracle_pipeline(frames, frame_base=lo)       # stands in for the kernels on this CPU box
    gathered = sharding.all_gather_histogram(torch.from_numpy(o["hist"]))
    ids = [int(r["frame"]) for f in range(len(frames)) for r in o["recs"][f, : o["n_out"][f]]]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), gathered=gathered, local=o["hist"], frame_ids=np.array(ids),
             n_out=o["n_out"])
    dist.destroy_process_group()


def test_two_rank_shards_and_histogram_allgather(tmp_path):
    """World size 2 over gloo: the union of the shards equals the single-process run and the
    all-gathered per-class histogram sums to the bincount over every emitted record (S7)."""
    import torch.multiprocessing as mp

    n_frames, world = 5, 2
    mp.spawn(_shard_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    spec = synthetic.SceneSpec(256, 144, 12, 2, 17, config_id=7)
    single = helpers.oracle_pipeline([synthetic.make_frame(spec, i) for i in range(n_frames)])
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert np.array_equal(parts[0]["gathered"], parts[1]["gathered"])         # every rank sees every rank
    assert np.array_equal(parts[0]["gathered"], np.stack([p["local"] for p in parts]))
    assert np.array_equal(parts[0]["gathered"].sum(axis=0), single["hist"])
    all_classes = np.concatenate([single["recs"][f, : single["n_out"][f]]["class_id"] for f in range(n_frames)])
    assert np.array_equal(single["hist"], np.bincount(all_classes, minlength=10))
    assert np.array_equal(np.concatenate([p["n_out"] for p in parts]), single["n_out"])
    want_ids = [f for f in range(n_frames) for _ in range(single["n_out"][f])]
    assert list(np.concatenate([p["frame_ids"] for p in parts])) == want_ids   # global frame ids, in order


def test_native_yolo_formatter_matches_python(libcspe_path):
    """f3: cspe_format_yolo_host is byte-identical to the Python formatter (no GPU needed)."""
    rng = np.random.default_rng(5)
    B, N = 5, 40
    recs = np.zeros((B, N), dtype=O.RECORD_DTYPE)
    n_out = rng.integers(0, N + 1, size=B).astype(np.int32)
    n_out[1] = 0
    recs["class_id"] = rng.integers(0, 10, size=(B, N))
    vals = rng.uniform(0, 1, size=(B, N, 4)).astype(np.float32)
    # rounding edge cases: exact ties at the 7th decimal, 0, 1, tiny and > 1 values
    vals[0, 0] = [0.0, 1.0, 0.5, 0.0000005]
    vals[0, 1] = [0.0000015, 0.9999995, 0.1234565, 2.5]
    vals[0, 2] = [np.float32(1) / 3, np.float32(2) / 3, 1e-7, 0.999999]
    vals[2, :, :] = (rng.integers(0, 2_000_001, size=(N, 4)) / 2_000_000).astype(np.float32)   # near-tie grid
    recs["yolo"] = vals
    buf, off = formats.yolo_text_batch(recs, n_out)
    for f in range(B):
        lines = formats.yolo_lines(recs[f, : n_out[f]])
        want = ("\n".join(lines) + "\n") if lines else ""
        assert bytes(buf[off[f]:off[f + 1]]).decode() == want, f
    buf3, off3 = formats.yolo_text_batch(recs, n_out, frames=3)
    assert off3.shape == (4,) and bytes(buf3) == bytes(buf[: off[3]])
    bad = recs.copy()
    bad["yolo"][0, 0, 0] = np.nan
    from constructionsceneposeestimation_b200 import _lib
    with pytest.raises(_lib.CspeError):
        formats.yolo_text_batch(bad, n_out)


# ------------------------------------------------------------------ f3: quality log (gcd.py:236-464)
def _replay_quality(events, depth_dict_of):
    from constructionsceneposeestimation_b200.quality import FrameQualityLog

    log = FrameQualityLog()
    for ev in events:
        log.frame_start(ev["frame"], np.asarray(ev["cam"]))
        if "retry" in ev:
            log.retry(ev["retry"])
        if "cloud" in ev:
            log.pointcloud(*ev["cloud"])
        if "rgb" in ev:
            log.rgb(*ev["rgb"])
        if "depth" in ev:
            log.depth(None if ev["depth"] is None else depth_dict_of(ev["depth"]), reason="annotator返回None或空")
        if "labels" in ev:
            log.labels(ev["labels"])
        log.frame_end(ev["ok"])
    return log


def test_quality_log_matches_reference_logger(tmp_path):
    """Same events -> the generation_summary.json the reference's DataQualityLogger wrote (golden)."""
    gold = json.loads((GOLDEN / "quality_log.json").read_text(encoding="utf-8"))
    cases = np.load(GOLDEN / "depth_stats.npz")
    log = _replay_quality(gold["events"], lambda name: O.depth_stats(cases[name]))
    log.log_dir = str(tmp_path / "logs")
    data = log.save_summary()
    assert data == gold["summary"]
    on_disk = json.loads((tmp_path / "logs" / "generation_summary.json").read_text(encoding="utf-8"))
    assert on_disk == gold["summary"]
    assert list(on_disk["statistics"]) == list(gold["summary"]["statistics"])       # key order too
    assert [f"{k}: {v} 次" for k, v in log.issue_counts().items()] == gold["issue_lines"]   # gcd.py:447-455
    # the readable report: the text the reference's own _generate_report returned (gcd.py:420-457), and its file
    assert log.report() == gold["report"]
    assert (tmp_path / "logs" / "generation_report.txt").read_text(encoding="utf-8") == gold["report"]
    log.save_summary(str(tmp_path / "logs" / "generation_summary_rank03.json"))
    assert (tmp_path / "logs" / "generation_report_rank03.txt").read_text(encoding="utf-8") == gold["report"]
    empty = __import__("constructionsceneposeestimation_b200.quality", fromlist=["FrameQualityLog"]).FrameQualityLog()
    assert "成功率: 0.0%" in empty.report() and empty.report().endswith("常见问题:\n")


def test_depth_quality_from_device_record_layout():
    """cspe_depth_stats_t -> the reference's depth dict (host conversion only; the kernel is a GPU test)."""
    from constructionsceneposeestimation_b200 import _lib
    from constructionsceneposeestimation_b200.quality import depth_quality_from_stats

    cases = np.load(GOLDEN / "depth_stats.npz")
    want = json.loads(str(cases["results"]))
    for name, ref in want.items():
        d = cases[name]
        ok = np.isfinite(d) & (d > 0)
        st = np.zeros(1, dtype=_lib.DEPTH_STATS_DTYPE)[0]
        st["valid_pixels"], st["zero_pixels"], st["inf_pixels"], st["total_pixels"] = ok.sum(), (d == 0).sum(), np.isinf(d).sum(), d.size
        if ok.any():
            st["depth_min"], st["depth_max"], st["depth_sum"] = d[ok].min(), d[ok].max(), d[ok].astype(np.float64).sum()
        got = depth_quality_from_stats(st)
        assert list(got) == list(ref)
        for k in ref:
            if k == "depth_mean":
                assert abs(got[k] - ref[k]) <= 1e-5 * max(1.0, abs(ref[k]))
            else:
                assert got[k] == ref[k], (name, k)


# ------------------------------------------------------------------ f3: native label JSON (gcd.py:608-613, 2056-2064)
def _random_records(rng, n, num_slots):
    from constructionsceneposeestimation_b200 import _lib

    recs = np.zeros(n, dtype=_lib.RECORD_DTYPE)
    recs["inst_idx"] = np.sort(rng.choice(num_slots, n, replace=False))
    recs["class_id"] = rng.integers(0, 10, n)
    recs["count"] = rng.integers(0, 10 ** 6, n)
    for k in ("x_min", "y_min", "x_max", "y_max"):
        recs[k] = rng.integers(-1, 4000, n)
    recs["loose"] = rng.integers(-5, 4000, (n, 4))
    recs["flags"] = rng.integers(0, 16, n)
    for k in ("occlusion", "fill", "truncation"):
        recs[k] = rng.uniform(size=n).astype(np.float32)
    # magnitudes across Python's fixed / exponent repr switch (1e-4, 1e16), integers, negative zero, specials
    mag = 10.0 ** rng.integers(-12, 22, (n, 16))
    recs["pose"] = rng.normal(size=(n, 16)) * mag
    recs["uv"] = rng.uniform(-100, 4000, (n, 8, 2))
    recs["z"] = np.round(rng.uniform(-5, 300, (n, 8)), 1)
    if n > 8:
        recs["pose"][0, :8] = [0.0, -0.0, 1e16, 9999999999999998.0, 1e-4, 0.00009999, 5e-324, 1.7976931348623157e308]
        recs["pose"][1, 7], recs["pose"][1, 14], recs["uv"][2, 3, 0], recs["z"][3, 7] = np.nan, np.inf, -np.inf, np.nan
        recs["pose"][4, :3] = [100.0, 123456.0, 1e15]
        recs["occlusion"][5], recs["truncation"][6], recs["fill"][7] = np.nan, np.inf, 1.0
    return recs


def test_native_label_json_matches_python_json_dump(libcspe_path):
    """cspe_format_label_json_host == json.dumps(reference_label(...), indent=2, ensure_ascii=False), byte for byte."""
    from constructionsceneposeestimation_b200.classes import SceneObject

    rng = np.random.default_rng(2026)
    names = ["fence", "tree", 'we"ird\\name', "吊车\ttab", "ctl\x01\n", "human"]
    objects = [SceneObject(inst_idx=i, class_id=i % 10, class_name=names[i % len(names)],
                           prim_path=f"/World/锥/obj_{i:03d}/\"q\"") for i in range(40)]
    pose = [1.5, -2.0, 3.25, 0.1, 0.2, 0.3, 0.9]
    params = {"horizontal_aperture": 25.0, "vertical_aperture": 14.0625, "focal_length": 12.0, "width": 1280, "height": 720}
    for n in (0, 1, 17, 40):
        recs = _random_records(rng, n, 40)
        want = json.dumps(formats.reference_label(7, pose, params, 720, 1280, recs, objects), indent=2, ensure_ascii=False)
        assert formats.label_json_bytes(7, pose, params, 720, 1280, recs, objects) == want.encode("utf-8"), n
    # keypoint blocks, unusual camera dicts and non-finite camera poses
    recs = _random_records(rng, 12, 40)
    P, J = 3, 5
    kp = rng.uniform(-50, 2000, (P, J, 2))
    vis = rng.integers(0, 3, (P, J)).astype(np.uint8)
    kp[0, 1, 0], kp[2, 4, 1] = np.nan, np.inf
    person_slots = [int(recs["inst_idx"][2]), -1, int(recs["inst_idx"][9])]
    blocks = {s: formats.coco_keypoint_block(kp[p], vis[p]) for p, s in enumerate(person_slots) if s >= 0}
    for prm, cam in ((params, pose), ({}, [np.nan, np.inf, -np.inf, 0, 0, 0, 1]), ({"k": [1, {"a": None}], "ü": "é"}, pose)):
        want = json.dumps(formats.reference_label(123456, cam, prm, 2160, 3840, recs, objects, blocks), indent=2,
                          ensure_ascii=False)
        got = formats.label_json_bytes(123456, cam, prm, 2160, 3840, recs, objects, None, kp, vis, person_slots)
        assert got == want.encode("utf-8")
    with pytest.raises(Exception):
        bad = recs.copy()
        bad["inst_idx"][0] = 99
        formats.label_json_bytes(0, pose, params, 720, 1280, bad, objects)


def test_native_label_json_reproduces_reference_file():
    """The golden text was written by the reference's own save_label_json (gcd.py:608-613)."""
    from constructionsceneposeestimation_b200 import _lib
    from constructionsceneposeestimation_b200.classes import SceneObject

    gold = json.loads((GOLDEN / "label_schema.json").read_text(encoding="utf-8"))
    lab, obj = gold["label"], gold["label"]["objects"][0]
    rec = np.zeros(1, dtype=_lib.RECORD_DTYPE)
    rec["inst_idx"], rec["class_id"], rec["flags"] = obj["inst_idx"], obj["class_id"], 15
    rec["pose"][0, 7:10], rec["pose"][0, 10:13], rec["pose"][0, 13:16] = obj["center"], obj["size"], obj["rotation"]
    objects = [SceneObject(inst_idx=0, class_id=obj["class_id"], class_name=obj["class_name"], prim_path=obj["prim_path"])]
    ours = json.loads(formats.label_json_bytes(lab["frame_id"], lab["camera_pose"], lab["camera_params"],
                                               *lab["instance_mask_shape"], rec, objects).decode("utf-8"))
    # reference keys, in the reference's order, with the reference's values; our extra keys come after
    assert list(ours)[:7] == list(lab)
    for k in lab:
        if k != "objects":
            assert ours[k] == lab[k], k
    assert list(ours["objects"][0])[:7] == list(obj) and all(ours["objects"][0][k] == obj[k] for k in obj)
    # and the layout (indent=2, one element per line) is the reference file's: strip our added keys and compare text
    trimmed = dict(ours)
    trimmed["objects"] = [{k: ours["objects"][0][k] for k in obj}]
    assert json.dumps(trimmed, indent=2, ensure_ascii=False) == gold["text"]


def test_native_coco_annotations_match_python_json_dump(libcspe_path):
    """cspe_format_coco_host == the inside of json.dumps(coco_annotations(...) of every frame)."""
    from constructionsceneposeestimation_b200 import _lib

    rng = np.random.default_rng(77)
    B, N, P, J = 4, 24, 3, 5
    recs = np.zeros((B, N), dtype=_lib.RECORD_DTYPE)
    n_out = np.array([24, 0, 7, 13], dtype=np.int32)
    for f in range(B):
        r = _random_records(rng, int(n_out[f]), N) if n_out[f] else recs[f, :0]
        recs[f, : n_out[f]] = r
    recs["count"][0, 3] = 0          # empty box -> [0, 0, 0, 0]
    image_ids = [100, 101, 205, 206]
    kp = rng.uniform(-50, 2000, (B, P, J, 2))
    vis = rng.integers(0, 3, (B, P, J)).astype(np.uint8)
    kp[0, 1, 2, 0], kp[3, 0, 4, 1] = np.nan, np.inf
    person_slots = [[int(recs["inst_idx"][f, 0]) if n_out[f] else -1, -1, int(recs["inst_idx"][f, 2]) if n_out[f] > 2 else -1]
                    for f in range(B)]
    for with_kp in (False, True):
        anns = []
        for f in range(B):
            blocks = None
            if with_kp:
                blocks = {s: formats.coco_keypoint_block(kp[f, p], vis[f, p]) for p, s in enumerate(person_slots[f]) if s >= 0}
            anns += formats.coco_annotations(recs[f, : n_out[f]], image_ids[f], len(anns) + 11, blocks)
        want = json.dumps(anns)
        text, count = formats.coco_annotations_text(recs, n_out, image_ids, 11, *((kp, vis, person_slots) if with_kp else ()))
        assert count == len(anns) == int(n_out.sum())
        assert b"[" + text + b"]" == want.encode("ascii")
    text, count = formats.coco_annotations_text(recs, np.zeros(B, dtype=np.int32), image_ids, 1)
    assert text == b"" and count == 0


def test_tables_cache_key_accepts_replicator_label_shapes():
    """idToLabels as Replicator hands it over: string keys, prim-path strings or {"class": ...} dicts (gcd.py:1826-1837)."""
    from constructionsceneposeestimation_b200.writer import tables_cache_key

    labels = {"0": "BACKGROUND", "1": "UNLABELLED", "2": "/World/a/mesh", 3: {"class": "fence"}, 4: {"primPath": "/World/b"}, 5: None}
    key = tables_cache_key(["/World/a/mesh", "/World/b"], labels)
    assert hash(key) == hash(tables_cache_key(["/World/a/mesh", "/World/b"], dict(labels)))
    assert key != tables_cache_key(["/World/a/mesh", "/World/b"], {**labels, "2": "/World/c/mesh"})
    assert key[1][3] == ("3", None) and key[1][4] == ("4", "/World/b")


# ------------------------------------------------------------------ Replicator-native payloads (boundary)
def test_replicator_camera_params_and_pose_convention():
    """cameraViewTransform -> (t, Rcw, quaternion) exactly as get_obj_pose takes them from the camera prim
    (gcd.py:596-605): the inverse view matrix is the prim's local-to-world matrix; scipy's from_matrix().as_quat()
    is restated in camera.matrix_to_quat_xyzw and checked against scipy here."""
    from scipy.spatial.transform import Rotation
    from constructionsceneposeestimation_b200 import camera, synthetic
    rng = np.random.default_rng(3)
    for _ in range(200):
        m = Rotation.random(random_state=rng).as_matrix()
        assert np.allclose(camera.matrix_to_quat_xyzw(m), Rotation.from_matrix(m).as_quat(), atol=1e-15)
    fr = synthetic.make_frame(synthetic.CONFIGS["c1"], 3)
    pose, p = np.asarray(fr["camera_pose"]), fr["camera_params"]
    rcw = camera.quat_xyzw_to_matrix(pose[3:])
    m = np.eye(4)
    m[:3, :3], m[3, :3] = rcw.T, pose[:3]
    native = {"cameraViewTransform": np.linalg.inv(m).reshape(-1), "cameraFocalLength": p["focal_length"],
              "cameraAperture": [p["horizontal_aperture"], p["vertical_aperture"]],
              "renderProductResolution": [p["width"], p["height"]], "cameraNearFar": [0.1, 1000.0]}
    assert camera.is_replicator_camera_params(native) and not camera.is_replicator_camera_params(p)
    pose7, params, clip = camera.from_replicator_camera_params(native)
    assert params == pytest.approx(p) and clip == (0.1, 1000.0)
    want = camera.pack_camera(pose, p)
    got = camera.pack_camera(pose7, params)
    assert np.allclose(got, want, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        camera.from_replicator_camera_params({"cameraViewTransform": np.eye(4).reshape(-1)})
    # image size from the mask when the payload has no resolution
    _, params2, _ = camera.from_replicator_camera_params({"cameraViewTransform": np.eye(4).reshape(-1)}, 1920, 1080)
    assert (params2["width"], params2["height"]) == (1920, 1080)
    assert params2["horizontal_aperture"] == camera.DEFAULT_HORIZONTAL_APERTURE


def test_render_product_keys_and_skeleton_shapes():
    import json
    from constructionsceneposeestimation_b200.writer import _annotator_name, skeleton_joints, split_render_products
    assert _annotator_name("instance_segmentation-RenderProduct_Replicator") == ("instance_segmentation", "RenderProduct_Replicator")
    assert _annotator_name("bounding_box_3d_fast") == ("bounding_box_3d", None)
    assert _annotator_name("LdrColor-rp") == ("rgb", "rp")
    assert _annotator_name("trigger_outputs") == (None, None)
    one = split_render_products({"instance_segmentation": 1, "frame_id": 4})
    assert one == [{"instance_segmentation": 1, "frame_id": 4}]
    rig = split_render_products({"instance_segmentation-b": "mb", "instance_segmentation-a": "ma", "camera_params-a": "ca",
                                 "camera_params-b": "cb", "frame_id": 2, "trigger_outputs": {}})
    assert [fr["render_product"] for fr in rig] == ["a", "b"]
    assert rig[0]["instance_segmentation"] == "ma" and rig[1]["camera_params"] == "cb" and rig[1]["frame_id"] == 2
    j = np.arange(2 * 5 * 3, dtype=np.float32).reshape(2, 5, 3)
    for payload in (j, {"globalTranslations": j}, {"data": {"globalTranslations": j.tolist()}},
                    [{"skelPath": "/a", "globalTranslations": j[0]}, {"skelPath": "/b", "globalTranslations": j[1].tolist()}],
                    json.dumps([{"globalTranslations": j[0].tolist()}, {"globalTranslations": j[1].tolist()}]),
                    {"skeletonData": [{"globalTranslations": j[0]}, {"globalTranslations": j[1]}]}):
        got = skeleton_joints(payload)
        assert got is not None and got.dtype == np.float32 and np.array_equal(got, j)
    assert np.array_equal(skeleton_joints(j[0]), j[:1])
    for bad in (None, "not json", {"skelPath": "/a"}, [{"globalTranslations": j[0]}, {"globalTranslations": j[1, :3]}],
                np.zeros((4, 2))):
        assert skeleton_joints(bad) is None


def test_record_index_marks_stand_in_records():
    """first_mesh fallback: a mesh record standing in for a multi-mesh object carries RECORD_APPROX_BIT (K2 turns it into
    OBJ_APPROX_RECORD); crane parts (the reference's own rule, gcd.py:1953-1975), single-mesh objects and objects with a
    record of their own do not; a non-numeric idToLabels key is skipped, not fatal."""
    from constructionsceneposeestimation_b200 import classes, synthetic
    res = classes.ObjectRootResolver()
    fence = synthetic.FENCE_PREFIX + "03"
    paths = [fence + "/Mesh_0", fence + "/Mesh_1", "/World/Tree/Tree", "/World/Tree/Tree/trunk",
             "/World/GroundPlane/Cone001/Cone001", synthetic.CRANE_ROOT + "/S104GG03A_SW/part_0",
             synthetic.CRANE_ROOT + "/S104GG03A_SW/part_1"]
    objs = classes.aggregate_objects(paths, res)
    idx = classes.record_index_for(objs, paths, "first_mesh")
    by_name = {o.class_name: i for o, i in zip(objs, idx)}
    assert by_name["fence"] == (0 | classes.RECORD_APPROX_BIT)
    assert by_name["tree"] == 2 and by_name["trafficcone"] == 4
    crane = [i for o, i in zip(objs, idx) if "#" in o.prim_path]
    assert crane and all(i >= 0 and not (i & classes.RECORD_APPROX_BIT) for i in crane)
    assert classes.record_index_for(objs, paths, "first_mesh", mark_approx=False)[0] == 0
    assert classes.record_index_for(objs, paths, "reference")[0] == -1
    m = classes.id_to_slot({"7": paths[0], "x9": paths[1], 8: paths[2]}, objs, res)
    assert m == {7: 0, 8: objs[[o.class_name for o in objs].index("tree")].inst_idx}


def test_native_batch_file_writer(tmp_path):
    """cspe_write_files_host: label_%06d naming of gcd.py:2071, one file per row of a strided text buffer."""
    from constructionsceneposeestimation_b200 import _lib, formats
    data = np.zeros((4, 32), dtype=np.uint8)
    texts = [b"0 0.5 0.5 0.1 0.1\n", b"", b"3 0.25 0.75 0.5 0.5\n1 0 0 1 1\n"[:32], b"x" * 32]
    for j, t in enumerate(texts):
        data[j, : len(t)] = np.frombuffer(t, dtype=np.uint8)
    sizes = np.array([len(t) for t in texts], dtype=np.int32)
    assert formats.write_files(str(tmp_path), "label_", ".txt", 41, data, sizes) == int(sizes.sum())
    for j, t in enumerate(texts):
        assert (tmp_path / f"label_{41 + j:06d}.txt").read_bytes() == t
    formats.write_files(str(tmp_path), "label_", ".txt", 41, data, np.array([3, 0, 0, 0], dtype=np.int32), count=1)
    assert (tmp_path / "label_000041.txt").read_bytes() == texts[0][:3]          # truncated, not appended
    with pytest.raises(_lib.CspeError):
        formats.write_files(str(tmp_path / "missing_dir"), "label_", ".txt", 0, data, sizes)
    with pytest.raises(_lib.CspeError):
        formats.write_files(str(tmp_path), "label_", ".txt", 0, data, np.array([33, 0, 0, 0], dtype=np.int32))


def test_concat_rows_and_coco_ratio_rule():
    """cspe_concat_rows_host (strided rows back to back) and the 6-decimal rule of the COCO ratio fields: the native host
    formatter prints them exactly as Python prints round(x, 6) — fixed notation from 1e-4 up, exponent form below,
    trailing zeros dropped (csrc/repr6.h is shared with the device formatter)."""
    import json
    from constructionsceneposeestimation_b200 import _lib, formats
    data = np.frombuffer(b"abcdefgh" + b"ij------" + b"--------" + b"klmnopqr", dtype=np.uint8).reshape(4, 8).copy()
    assert formats.concat_rows(data, np.array([8, 2, 0, 5], dtype=np.int32)) == b"abcdefghijklmno"
    assert formats.concat_rows(data, np.array([8, 2, 0, 5], dtype=np.int32), count=2) == b"abcdefghij"
    with pytest.raises(_lib.CspeError):
        formats.concat_rows(data, np.array([9, 0, 0, 0], dtype=np.int32))
    vals = np.array([0.0, 1.0, 0.5, 5e-7, 4.9e-7, 1.2e-5, 9.95e-5, 9.96e-5, 1e-4, 1e-6, 0.9999995, 0.99999994, 1e-30, 0.1234565,
                     0.000123, 0.25, 3.3e-5, 5.05e-5, 0.1, 0.7, 2.5e-6], dtype=np.float32)
    rng = np.random.default_rng(4)
    vals = np.concatenate([vals, rng.random(500, dtype=np.float32), (rng.integers(0, 2 ** 21, 500) / np.float32(2 ** 21)).astype(np.float32),
                           rng.random(200, dtype=np.float32) * np.float32(2e-4)])
    recs = np.zeros((1, len(vals)), dtype=_lib.RECORD_DTYPE)
    recs["count"], recs["x_max"], recs["y_max"] = 7, 3, 4
    recs["occlusion"][0] = vals
    recs["truncation"][0] = vals[::-1]
    n = np.array([len(vals)], dtype=np.int32)
    text, count = formats.coco_annotations_text(recs, n, [12], 1)
    py = formats.coco_annotations(recs[0], 12, 1)
    assert count == len(vals) and text == json.dumps(py)[1:-1].encode()
    assert all(a["occlusion"] == round(float(v), 6) for a, v in zip(py, vals))
    assert b'e-06, ' in text and b'"occlusion": 1.2e-05' in text and b'"occlusion": 0.0001,' in text


def test_union_record_fallback_rules():
    """record_fallback="union": a multi-mesh object whose root has no record points behind the frame's own records
    (index len(primPaths) + u, APPROX bit set) and union_members lists the records of all its meshes; crane parts keep
    the reference's first-mesh rule (gcd.py:1953-1975), single-mesh objects their only mesh, objects with a record of
    their own that record; pack_union pads the batch's CSR tables."""
    from constructionsceneposeestimation_b200 import classes
    res = classes.ObjectRootResolver()
    fa, fb_root = synthetic.FENCE_PREFIX + "03", synthetic.FENCE_PREFIX + "04"
    paths = [fa + "/Mesh_0", fa + "/Mesh_1", fa + "/Mesh_2",                                  # no root record: union
             fb_root, fb_root + "/Mesh_0", fb_root + "/Mesh_1",                               # root record wins
             "/World/GroundPlane/Cone001_01/Cone001",                                         # single mesh: exact
             "/World/Tree/Tree_02/leaves", "/World/Tree/Tree_02/trunk"]                       # union
    objs = classes.aggregate_objects(paths, res)
    by_path = {o.prim_path: i for i, o in enumerate(objs)}
    idx = classes.record_index_for(objs, paths, "union")
    plan = classes.union_members(objs, paths)
    assert len(objs) == 4
    assert [s for s, _ in plan] == sorted(by_path[p] for p in by_path if p == fa or p.endswith("Tree_02"))
    for u, (slot, mem) in enumerate(plan):
        assert idx[slot] == (len(paths) + u) | classes.RECORD_APPROX_BIT
        assert [paths[m].startswith(objs[slot].prim_path + "/") for m in mem] == [True] * len(mem)
    fb = by_path[fb_root]
    cone = next(i for p, i in by_path.items() if "Cone" in p)
    assert idx[fb] == 3 and idx[cone] == 6
    # the other fallbacks on the same scene
    first = classes.record_index_for(objs, paths, "first_mesh")
    assert first[plan[0][0]] == 0 | classes.RECORD_APPROX_BIT and classes.record_index_for(objs, paths, "reference")[plan[0][0]] == -1
    off, mem, U = classes.pack_union([plan, [], plan[:1]])
    assert U == 2 and off.shape == (3, 3) and mem.shape == (3, 5)
    assert off.tolist() == [[0, 3, 5], [0, 0, 0], [0, 3, 3]] and mem[0].tolist() == [0, 1, 2, 7, 8] and mem[1].tolist() == [-1] * 5
    off0, mem0, U0 = classes.pack_union([[], []])
    assert U0 == 0 and off0.shape == (2, 1) and mem0.shape == (2, 1)


def test_union_records_oracle_is_the_aligned_range():
    """oracle.union_records against a direct statement: every corner of every member, transformed, min / max."""
    rng = np.random.default_rng(3)
    recs = np.zeros((1, 5), dtype=O.BBOX3D_DTYPE)
    pts = []
    for r in range(3):
        lo, hi = rng.uniform(-2, 0, 3).astype(np.float32), rng.uniform(0.5, 2, 3).astype(np.float32)
        recs[0, r]["x_min"], recs[0, r]["y_min"], recs[0, r]["z_min"] = lo
        recs[0, r]["x_max"], recs[0, r]["y_max"], recs[0, r]["z_max"] = hi
        m = np.eye(4, dtype=np.float32)
        m[3, :3] = rng.uniform(-5, 5, 3)
        m[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        recs[0, r]["transform"] = m
        for k in range(8):
            c = np.array([hi[0] if k & 1 else lo[0], hi[1] if k & 2 else lo[1], hi[2] if k & 4 else lo[2], 1.0])
            pts.append(c @ m.astype(np.float64))
    out = O.union_records(recs, 3, np.array([0, 3, 3], dtype=np.int32), np.array([0, 1, 2], dtype=np.int32))
    pts = np.array(pts)[:, :3]
    got = out[0, 3]
    assert np.allclose([got["x_min"], got["y_min"], got["z_min"]], pts.min(0), rtol=1e-6, atol=1e-6)
    assert np.allclose([got["x_max"], got["y_max"], got["z_max"]], pts.max(0), rtol=1e-6, atol=1e-6)
    assert np.array_equal(got["transform"], np.eye(4)) and np.isnan(out[0, 4]["x_min"])   # object 1 has no member
    assert out[0, :3].tobytes() == recs[0, :3].tobytes()


def test_write_coco_file_accepts_bytes_and_array_chunks(tmp_path):
    """write_coco_file assembles the file from pre-formatted chunks: bytes or uint8 arrays (views of the pinned result
    buffer of a sweep), empty chunks skipped, separators only where the chunks do not carry their own."""
    import json
    anns = [{"id": 1, "image_id": 0, "category_id": 2, "bbox": [1, 2, 3, 4], "area": 5, "iscrowd": 0},
            {"id": 2, "image_id": 0, "category_id": 0, "bbox": [0, 0, 1, 1], "area": 1, "iscrowd": 0},
            {"id": 3, "image_id": 1, "category_id": 1, "bbox": [9, 9, 2, 2], "area": 3, "iscrowd": 0}]
    chunk = lambda a: json.dumps(a)[1:-1].encode()
    images = [formats.coco_image(0, 64, 48, "rgb_000000.png"), formats.coco_image(1, 64, 48, "rgb_000001.png")]
    want = json.dumps({"images": images, "annotations": anns, "categories": formats.coco_categories()})
    # host-formatter style: one chunk per batch, joined with ", "
    formats.write_coco_file(tmp_path / "a.json", images, [chunk(anns[:2]), b"", chunk(anns[2:])])
    assert (tmp_path / "a.json").read_text() == want
    # device style: chunks carry their own ", " and arrive as uint8 arrays; images as pre-formatted text
    dev = [np.frombuffer(chunk(anns[:1]), dtype=np.uint8), np.zeros(0, dtype=np.uint8),
           np.frombuffer(b", " + chunk(anns[1:]), dtype=np.uint8)]
    formats.write_coco_file(tmp_path / "b.json", [formats.coco_images_text(range(0, 2), 64, 48)], dev, joined=True)
    assert (tmp_path / "b.json").read_text() == want
    assert formats.coco_images_text(range(0, 2), 64, 48) == formats.coco_images_text([0, 1], 64, 48)


def _host_only_writer(record_fallback="first_mesh"):
    """A ConstructionLabelWriter without a device: just what _host_tables touches (the constructor needs CUDA and the
    built library — this box has no GPU).  Ordinary memory stands in for the pinned blocks; test scaffolding only."""
    import torch
    from constructionsceneposeestimation_b200.writer import ConstructionLabelWriter
    w = ConstructionLabelWriter.__new__(ConstructionLabelWriter)
    w.device = torch.device("cpu")
    w.record_fallback = record_fallback
    w.resolver = classes.ObjectRootResolver(None, split_people=True)
    w.near, w.far = 0.5, 250.0
    w.max_lut_entries = 1 << 28
    w._tables_cache = {}
    w._warned_empty_lut = False
    w._next_frame_id = 0

    def take(kind, shape, dtype, owned):
        t = torch.zeros(shape, dtype=dtype)
        owned.append(((kind, shape, dtype), t))
        return t

    w._take_pinned = take
    return w


@pytest.mark.parametrize("fallback", ["first_mesh", "union"])
def test_writer_host_tables_match_the_helpers(fallback):
    """ConstructionLabelWriter._host_tables (the numpy-only half of annotate_batch) lays out the same LUT / slot /
    record / camera tables as the test helpers build from classes.py + the oracle's camera packing — for a list of
    frame dicts, for the same frames cut from ONE stacked dict, and for a batch that shows one scene (shared tables)."""
    from constructionsceneposeestimation_b200 import writer as W
    spec = synthetic.SceneSpec(640, 360, 24, 3, 17, config_id=31)
    frames = synthetic.make_batch(spec, 4)
    want = helpers.host_tables(frames, fallback=fallback)
    lut_w, obj_w, cls_w, rec_w, cam_w, objs_w = want
    R0 = helpers.host_tables.union[0]
    records_in = helpers.host_tables.union[3]      # before the union pass (the device builds those records)
    stacked = {
        "instance_segmentation": {"data": np.stack([f["instance_segmentation"]["data"] for f in frames]),
                                  "info": [f["instance_segmentation"]["info"] for f in frames]},
        "bounding_box_3d": {"data": [f["bounding_box_3d"]["data"] for f in frames],
                            "info": [f["bounding_box_3d"]["info"] for f in frames]},
        "camera_pose": np.asarray([f["camera_pose"] for f in frames]),
        "camera_params": [f["camera_params"] for f in frames],
        "frame_id": 40,
    }
    for form in ("list", "stacked"):
        w = _host_only_writer(fallback)
        if form == "list":
            fr = [w._normalise(f) for f in frames]
        else:
            cut = W._unstack(stacked)
            assert cut.canonical
            fr = [w._normalise_camera(f) for f in cut]
        hb = w._host_tables(fr, 360, 640, [])
        blk = hb.block
        assert not hb.same_tables and hb.N == obj_w.shape[1] and hb.R0 == R0
        assert hb.frame_ids == ([0, 1, 2, 3] if form == "list" else [40, 41, 42, 43]) and hb.contiguous_ids
        L = lut_w.shape[1]
        assert np.array_equal(blk.lut[:, :L], lut_w) and (blk.lut[:, L:] == -1).all()
        assert np.array_equal(blk.obj_record, obj_w) and np.array_equal(blk.slot_class, cls_w)
        assert np.array_equal(blk.cam, cam_w)
        got_recs = blk.records.reshape(4, -1).view(O.BBOX3D_DTYPE).reshape(4, -1)
        assert got_recs[:, :R0].tobytes() == records_in[:, :R0].tobytes()
        assert [[o.prim_path for o in t.objects] for t in hb.tables] == [[o.prim_path for o in objs] for objs in objs_w]
        if fallback == "union":
            off, mem = helpers.host_tables.union[1], helpers.host_tables.union[2]
            assert hb.U == off.shape[1] - 1 > 0
            assert np.array_equal(blk.union_offsets, off) and np.array_equal(blk.union_members, mem)
        else:
            assert hb.U == 0
    # one scene for the whole batch: ONE LUT row, broadcast slot tables; frames without records or ids
    w = _host_only_writer(fallback)
    same = [dict(frames[0], frame_id=7 + 2 * i) for i in range(3)]
    same[1] = dict(same[1], bounding_box_3d={"data": None, "info": frames[0]["bounding_box_3d"]["info"]})
    hb = w._host_tables([w._normalise(f) for f in same], 360, 640, [])
    assert hb.same_tables and hb.block.lut.shape[0] == 1 and not hb.contiguous_ids and hb.frame_ids == [7, 9, 11]
    assert np.array_equal(hb.block.lut[0, :lut_w.shape[1]], lut_w[0])
    assert np.array_equal(hb.block.slot_class, np.repeat(cls_w[:1], 3, axis=0))
    assert (hb.block.obj_record[1] == -1).all() and np.array_equal(hb.block.obj_record[0], obj_w[0])   # no records -> no boxes
    assert (hb.block.records[1] == 0).all()


def test_union_plan_property():
    """Random scenes (roots with / without a record of their own, 1-4 meshes, meshes without records, a crane part):
    record_index_for("union") and union_members agree object by object, members are exactly the records of the
    object's mesh paths, pack_union round-trips, and the other fallbacks are untouched by the union rule."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(st.booleans(), st.integers(1, 4), st.integers(0, 15)), min_size=1, max_size=12), st.booleans())
    def run(objs_spec, with_crane):
        paths = []
        for i, (own, n_mesh, drop_mask) in enumerate(objs_spec):
            root = f"{synthetic.FENCE_PREFIX}{i + 3:02d}"
            if own:
                paths.append(root)
            for m in range(n_mesh):
                if not (drop_mask >> m) & 1:            # a mesh without a bbox3d record simply is not listed
                    paths.append(f"{root}/Mesh_{m}")
        if with_crane:
            paths += [f"{synthetic.CRANE_ROOT}/{synthetic.CRANE_CHILDREN[0]}/part_0", f"{synthetic.CRANE_ROOT}/{synthetic.CRANE_CHILDREN[0]}/part_1"]
        if not paths:
            return
        res = classes.ObjectRootResolver()
        objs = classes.aggregate_objects(paths, res)
        idx = classes.record_index_for(objs, paths, "union")
        first = classes.record_index_for(objs, paths, "first_mesh")
        ref = classes.record_index_for(objs, paths, "reference")
        plan = dict(classes.union_members(objs, paths))
        u = 0
        for slot, o in enumerate(objs):
            own = o.actual_prim_path in paths
            multi = "#" not in o.prim_path and len(o.mesh_paths) > 1
            if own:
                assert idx[slot] == first[slot] == ref[slot] == paths.index(o.actual_prim_path) and slot not in plan
            elif multi:
                assert slot in plan and idx[slot] == (len(paths) + u) | classes.RECORD_APPROX_BIT
                assert plan[slot] == [paths.index(mp) for mp in o.mesh_paths] and ref[slot] == -1
                assert first[slot] == plan[slot][0] | classes.RECORD_APPROX_BIT
                u += 1
            else:   # single mesh or a crane part: the reference's own first-mesh rule, exact
                assert slot not in plan and idx[slot] == first[slot] == paths.index(o.mesh_paths[0])
        assert u == len(plan)
        off, mem, U = classes.pack_union([list(plan.items()), []])
        assert U == len(plan) and off[1].tolist() == [0] * (U + 1)
        for k, (slot, members) in enumerate(plan.items()):
            assert mem[0, off[0, k]: off[0, k + 1]].tolist() == members

    run()


def test_writer_host_tables_edge_cases():
    """_host_tables on the inputs the reference tolerates (gcd.py:1682, 1788, 1919, 2024-2027): scenes of different
    sizes in one batch (padding), a frame without bbox3d, records / primPaths of different lengths (cut to the
    shorter), no idToLabels entry that resolves (one warning), an id space too sparse for the dense LUT (clear error),
    default camera (script constants) and identity pose when the frame brings none."""
    import warnings as W
    spec_a = synthetic.SceneSpec(320, 180, 10, 2, 17, config_id=41)
    spec_b = synthetic.SceneSpec(320, 180, 18, 2, 17, config_id=42)
    fa, fb = synthetic.make_frame(spec_a, 0), synthetic.make_frame(spec_b, 1)
    w = _host_only_writer()
    hb = w._host_tables([w._normalise(fa), w._normalise(fb)], 180, 320, [])
    na, nb = len(hb.tables[0].objects), len(hb.tables[1].objects)
    assert na < nb == hb.N and not hb.same_tables
    assert (hb.block.obj_record[0, na:] == -1).all() and (hb.block.slot_class[0, na:] == -1).all()
    ra, rb = len(fa["bounding_box_3d"]["data"]), len(fb["bounding_box_3d"]["data"])
    assert hb.R0 == max(ra, rb) and (hb.block.records[0, ra:] == 0).all()
    # records longer than primPaths (and the other way round): both cut to the shorter, tables built from the cut paths
    long_recs = dict(fa, bounding_box_3d={"data": np.concatenate([fa["bounding_box_3d"]["data"]] * 2),
                                          "info": fa["bounding_box_3d"]["info"]})
    short_recs = dict(fa, bounding_box_3d={"data": fa["bounding_box_3d"]["data"][:5], "info": fa["bounding_box_3d"]["info"]})
    hb2 = w._host_tables([w._normalise(long_recs), w._normalise(short_recs)], 180, 320, [])
    assert hb2.R0 == ra and len(hb2.tables[0].objects) == na
    paths5 = fa["bounding_box_3d"]["info"]["primPaths"][:5]
    objs5 = classes.aggregate_objects(paths5, w.resolver)
    assert [o.prim_path for o in hb2.tables[1].objects] == [o.prim_path for o in objs5]
    assert (hb2.block.records[1, 5:] == 0).all()
    # no bbox3d at all, no camera: defaults of the script, identity pose, frame ids continue from the last call
    bare = {"instance_segmentation": fa["instance_segmentation"]}
    w._next_frame_id = 12
    hb3 = w._host_tables([w._normalise(bare), w._normalise(bare)], 180, 320, [])
    assert hb3.frame_ids == [12, 13] and hb3.N == 1 and (hb3.block.obj_record == -1).all() and hb3.same_tables
    cam = hb3.block.cam[0]
    assert cam[:3].tolist() == [0, 0, 0] and np.array_equal(cam[3:12].reshape(3, 3), np.eye(3))
    assert cam[18] == 320 and cam[19] == 180 and cam[16] == 0.5 and cam[17] == 250.0
    # idToLabels that resolves to nothing: one warning for the whole run
    unresolved = dict(fa, instance_segmentation={"data": fa["instance_segmentation"]["data"],
                                                 "info": {"idToLabels": {"7": "/World/Nothing/here"}}})
    with W.catch_warnings(record=True) as seen:
        W.simplefilter("always")
        w._host_tables([w._normalise(unresolved)], 180, 320, [])
        w._host_tables([w._normalise(unresolved)], 180, 320, [])
    assert len([x for x in seen if "idToLabels" in str(x.message)]) == 1
    # ... but not for a frame whose mask is missing anyway
    w2 = _host_only_writer()
    with W.catch_warnings(record=True) as seen:
        W.simplefilter("always")
        w2._host_tables([w2._normalise(unresolved)], 180, 320, [0])
    assert not seen
    # an id space too sparse for the dense id -> slot table
    w3 = _host_only_writer()
    w3.max_lut_entries = 1 << 10
    labels = dict(fa["instance_segmentation"]["info"]["idToLabels"])
    labels[str(1 << 20)] = fa["bounding_box_3d"]["info"]["primPaths"][0]     # a path that resolves to an object
    sparse = dict(fa, instance_segmentation={"data": fa["instance_segmentation"]["data"], "info": {"idToLabels": labels}})
    with pytest.raises(ValueError, match="id->slot table"):
        w3._host_tables([w3._normalise(sparse)], 180, 320, [])
