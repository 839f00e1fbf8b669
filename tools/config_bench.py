"""Secondary measurements for the BASELINE configs other than the bench's (c1, c3, c4) and the
"next" rows (fused depth statistics, point cloud, depth colormap).  One JSON line per case.
Usage (B200): python tools/config_bench.py [reps]
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cv2
import numpy as np
import torch

from constructionsceneposeestimation_b200 import _lib, camera, ops, synthetic
from constructionsceneposeestimation_b200.pipeline import LabelPipeline
from constructionsceneposeestimation_b200.sweep import build_host_tables

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda")
PEAK = 6454.3


def timed(fn, n=reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def out(**kw):
    print(json.dumps(kw), flush=True)


def build_pipeline(frames, B, use_graph=False):
    lut, obj_record, slot_class, records, cam, _ = build_host_tables(frames)
    lut = np.pad(lut, ((0, 0), (0, (-lut.shape[1]) % 4)), constant_values=-1)
    u = len(frames)
    rep = (B + u - 1) // u
    H, W = frames[0]["instance_segmentation"]["data"].shape
    N = obj_record.shape[1]
    pipe = LabelPipeline(B, H, W, N, records.shape[1], lut.shape[1], dev, use_graph=use_graph)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tile = lambda a: t(a).repeat((rep,) + (1,) * (a.ndim - 1))[:B]
    pipe.mask.copy_(tile(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)))
    pipe.lut.copy_(tile(lut))
    pipe.obj_record.copy_(tile(obj_record))
    pipe.slot_class.copy_(tile(slot_class))
    pipe.records_in.copy_(tile(records.view(np.uint8).reshape(u, records.shape[1], -1)))
    pipe.cam.copy_(tile(cam))
    return pipe, (H, W, N)


# ---- c1: single 1280x720 frame, ~20 instances: latency of one graph replay ----------------------
frames = synthetic.make_batch(synthetic.CONFIGS["c1"], 1)
for g in (True, False):
    pipe, (H, W, N) = build_pipeline(frames, 1, use_graph=g)
    ms = timed(pipe.run, n=200)
    out(case=f"c1 single 720p frame, 20 instances ({'CUDA-graph replay' if g else 'eager PDL chain'}, back-to-back)",
        ms=round(ms, 4), frames_per_s=round(1000 / ms, 1), note="3.7 MB fits L2: latency-bound, not a roofline case")

# ---- c4: 4-camera 3840x2160 rig, 500 instances: 16 rig-camera frames per GPU ----------------------
frames = synthetic.make_batch(synthetic.CONFIGS["c4"], 4)
pipe, (H, W, N) = build_pipeline(frames, 16)
ms = timed(pipe.run)
scan_ms = timed(lambda: ops.mask_scan(pipe.mask, pipe.lut, N, out=pipe.scan, accumulate=True))
out(case="c4 16 x 2160p frames, 500 instances", ms_per_step=round(ms, 4), frames_per_s=round(16000 / ms, 1),
    scan_ms=round(scan_ms, 4), scan_GBs=round(16 * H * W * 4 / scan_ms / 1e6, 1),
    scan_frac=round(16 * H * W * 4 / scan_ms / 1e6 / PEAK, 3))
del pipe

# ---- c3: people keypoint path, 50 people x J joints at 1080p, 64 frames ---------------------------
for J in (17, 101):
    spec = synthetic.SceneSpec(1920, 1080, 60, 50, J, config_id=3)
    frames = synthetic.make_batch(spec, 8)
    lut, obj_record, slot_class, records, cam, _ = build_host_tables(frames)
    lut = np.pad(lut, ((0, 0), (0, (-lut.shape[1]) % 4)), constant_values=-1)
    H, W = frames[0]["instance_segmentation"]["data"].shape
    N = obj_record.shape[1]
    pipe = LabelPipeline(64, H, W, N, records.shape[1], lut.shape[1], dev, use_graph=False, num_people=50, num_joints=J)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tile = lambda a: t(a).repeat((8,) + (1,) * (a.ndim - 1))
    pipe.mask.copy_(tile(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)))
    pipe.lut.copy_(tile(lut)); pipe.obj_record.copy_(tile(obj_record)); pipe.slot_class.copy_(tile(slot_class))
    pipe.records_in.copy_(tile(records.view(np.uint8).reshape(8, records.shape[1], -1))); pipe.cam.copy_(tile(cam))
    pipe.joints.copy_(tile(np.stack([f["skeleton_data"]["globalTranslations"] for f in frames])))
    pipe.depth.copy_(tile(np.stack([f["distance_to_image_plane"] for f in frames])))
    kp_ms = timed(lambda: ops.keypoints(pipe.joints, pipe.depth, pipe.cam, 0.15))
    step_ms = timed(pipe.run, n=50)
    out(case=f"c3 64 x 1080p, 50 people x {J} joints", keypoints_alone_ms=round(kp_ms, 4),
        joints_per_s=round(64 * 50 * J / kp_ms * 1e3), pipeline_step_ms=round(step_ms, 4),
        frames_per_s=round(64000 / step_ms, 1))
    del pipe

# ---- next rows on c2-shaped data ------------------------------------------------------------------
frames = synthetic.make_batch(synthetic.CONFIGS["c2"], 8)
lut, obj_record, *_ = build_host_tables(frames)
N = obj_record.shape[1]
mask = torch.from_numpy(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)).to(dev).repeat(8, 1, 1)
depth = torch.from_numpy(np.stack([f["distance_to_image_plane"] for f in frames])).to(dev).repeat(8, 1, 1)
lut_d = torch.from_numpy(lut).to(dev).repeat(8, 1)
B, H, W = mask.shape
scan_ms = timed(lambda: ops.mask_scan(mask, lut_d, N))
stats_ms = timed(lambda: ops.depth_stats(depth))
fused_ms = timed(lambda: ops.mask_scan_depth_stats(mask, depth, lut_d, N))
out(case="f2 depth statistics, 64 x 1080p", scan_ms=round(scan_ms, 4), depth_stats_alone_ms=round(stats_ms, 4),
    depth_stats_GBs=round(B * H * W * 4 / stats_ms / 1e6, 1), fused_scan_plus_stats_ms=round(fused_ms, 4),
    fused_GBs=round(2 * B * H * W * 4 / fused_ms / 1e6, 1), fused_frac=round(2 * B * H * W * 4 / fused_ms / 1e6 / PEAK, 3))
lut_bgr = torch.from_numpy(np.ascontiguousarray(cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(-1, 1), cv2.COLORMAP_JET).reshape(256, 3))).to(dev)
st = ops.depth_stats(depth)
cm_ms = timed(lambda: ops.depth_colormap(depth, lut_bgr, st))
out(case="f4 depth colormap, 64 x 1080p", ms=round(cm_ms, 4), GBs=round(B * H * W * 7 / cm_ms / 1e6, 1))
fr = frames[0]
cam = torch.from_numpy(camera.pack_camera(fr["camera_pose"], fr["camera_params"])).to(dev)
rgb = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device=dev)
d0 = depth[0].contiguous()
pts, n = ops.depth_to_pointcloud(d0, rgb, cam)
pc_ms = timed(lambda: ops.depth_to_pointcloud(d0, rgb, cam))
npts = int(n.item())
out(case="f1 depth -> point cloud, one 1080p frame", ms=round(pc_ms, 4), points=npts,
    GBs=round((2 * H * W * 4 + H * W * 4 + npts * 48) / pc_ms / 1e6, 1), frames_per_s=round(1000 / pc_ms, 1))
