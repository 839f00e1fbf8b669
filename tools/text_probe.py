"""One point-cloud + text formatting pass of a 1080p frame (for ncu): python tools/text_probe.py [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from constructionsceneposeestimation_b200 import camera, ops, synthetic
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fr = synthetic.make_frame(synthetic.CONFIGS["c2"], 0)
dev = torch.device("cuda")
cam = torch.from_numpy(camera.pack_camera(fr["camera_pose"], fr["camera_params"])).to(dev)
depth = torch.from_numpy(fr["distance_to_image_plane"]).to(dev)
rgb = torch.randint(0, 256, depth.shape + (4,), dtype=torch.uint8, device=dev)
for _ in range(reps):
    pts, n = ops.depth_to_pointcloud(depth, rgb, cam)
    text, nb, _ = ops.format_fixed6(pts, n_rows=n, header="x y z r g b")
    text2, nb2, _ = ops.format_fixed6(depth)
torch.cuda.synchronize()
print(int(n.item()), int(nb.item()), int(nb2.item()))
