"""Where the host time of ConstructionLabelWriter.annotate_batch goes when the annotators are device-resident
(cProfile, 64 x 1080p config-2 frames, stacked dict with a CUDA mask batch).

    python tools/writer_host_profile.py [reps]
"""
import cProfile
import io
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from constructionsceneposeestimation_b200 import synthetic
from constructionsceneposeestimation_b200.writer import ConstructionLabelWriter

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda")
uniq = 16
frames = synthetic.make_batch(synthetic.CONFIGS["c2"], uniq)
frames = [frames[i % uniq] for i in range(64)]
d_mask = torch.from_numpy(np.stack([f["instance_segmentation"]["data"] for f in frames]).view(np.int32)).to(dev)
batch = {
    "instance_segmentation": {"data": d_mask, "info": [fr["instance_segmentation"]["info"] for fr in frames]},
    "bounding_box_3d": {"data": [fr["bounding_box_3d"]["data"] for fr in frames],
                        "info": [fr["bounding_box_3d"]["info"] for fr in frames]},
    "camera_pose": np.asarray([fr["camera_pose"] for fr in frames], dtype=np.float64),
    "camera_params": [fr["camera_params"] for fr in frames],
    "frame_id": list(range(64)),
}
writer = ConstructionLabelWriter(None, device=dev, split_people=True)
for _ in range(3):
    writer.annotate_batch(batch).synchronize()


def run(n):
    prev = None
    for _ in range(n):
        cur = writer.annotate_batch(batch)
        if prev is not None:
            int(prev.n_out.sum())
        prev = cur
    int(prev.n_out.sum())


t0 = time.perf_counter()
run(reps)
dt = time.perf_counter() - t0
print(f"plain: {dt / reps * 1e3:.3f} ms per 64-frame batch = {64 * reps / dt:.0f} frames/s")
pr = cProfile.Profile()
pr.enable()
run(reps)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
