"""Where does the end-to-end (host arrays -> records on host) time go?"""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import synthetic
from constructionsceneposeestimation_b200.writer import ConstructionLabelWriter
B = 64
frames = synthetic.make_batch(synthetic.CONFIGS["c2"], 8)
H, W = frames[0]["instance_segmentation"]["data"].shape
mask_host = torch.empty((B, H, W), dtype=torch.int32, pin_memory=True)
host_frames = []
for i in range(B):
    fr = frames[i % 8]
    mask_host[i].copy_(torch.from_numpy(fr["instance_segmentation"]["data"].view(np.int32)))
    hf = {k: v for k, v in fr.items() if k not in ("skeleton_data", "distance_to_image_plane")}
    hf["instance_segmentation"] = {"data": mask_host[i], "info": fr["instance_segmentation"]["info"]}
    hf["frame_id"] = i
    host_frames.append(hf)
w = ConstructionLabelWriter(None, split_people=True)
for _ in range(3):
    w.annotate_batch(host_frames).synchronize()
def t(fn, n=5):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return round(best * 1e3, 3)
masks = [f["instance_segmentation"]["data"] for f in host_frames]
print(json.dumps({
    "annotate_batch_ms": t(lambda: w.annotate_batch(host_frames).synchronize()),
    "stack_to_device_ms": t(lambda: w._stack_to_device(masks, torch.int32)),
    "single_copy_ms": t(lambda: mask_host.to("cuda", non_blocking=True)),
    "host_only_tables_ms": t(lambda: [w.frame_tables(f["bounding_box_3d"]["info"]["primPaths"], f["instance_segmentation"]["info"]["idToLabels"]) for f in host_frames]),
}))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): w.annotate_batch(host_frames).synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
