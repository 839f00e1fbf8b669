"""e2e of one 64 x 1080p batch through the writer: list of per-frame pinned arrays vs one stacked pinned array.
python tools/e2e_stacked_probe.py"""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import synthetic
from constructionsceneposeestimation_b200.writer import ConstructionLabelWriter, _unstack
B = 64
uniq = synthetic.make_batch(synthetic.CONFIGS["c2"], 8)
H, W = uniq[0]["instance_segmentation"]["data"].shape
mask_host = torch.empty((B, H, W), dtype=torch.int32, pin_memory=True)
for i in range(B):
    mask_host[i].copy_(torch.from_numpy(uniq[i % 8]["instance_segmentation"]["data"].view(np.int32)))
frames = []
for i in range(B):
    fr = dict(uniq[i % 8]); fr.pop("skeleton_data", None); fr.pop("distance_to_image_plane", None)
    fr["instance_segmentation"] = {"data": mask_host[i], "info": fr["instance_segmentation"]["info"]}; fr["frame_id"] = i
    frames.append(fr)
stacked = {"instance_segmentation": {"data": mask_host, "info": [f["instance_segmentation"]["info"] for f in frames]},
           "bounding_box_3d": {"data": [f["bounding_box_3d"]["data"] for f in frames], "info": [f["bounding_box_3d"]["info"] for f in frames]},
           "camera_pose": np.asarray([f["camera_pose"] for f in frames]), "camera_params": [f["camera_params"] for f in frames], "frame_id": 0}
w = ConstructionLabelWriter(None, split_people=True)
def run(make, steps=10):
    for _ in range(2): w.annotate_batch(make()).synchronize()
    torch.cuda.synchronize(); t0 = time.perf_counter(); prev = None
    for _ in range(steps):
        cur = w.annotate_batch(make())
        if prev is not None: prev.n_out.sum()
        prev = cur
    prev.n_out.sum(); return (time.perf_counter() - t0) / steps
for _ in range(2):
    a = run(lambda: frames); b = run(lambda: _unstack(stacked))
    print(json.dumps({"list_ms": round(a * 1e3, 3), "list_fps": round(B / a), "stacked_ms": round(b * 1e3, 3), "stacked_fps": round(B / b)}))
