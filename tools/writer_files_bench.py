"""End-to-end writer throughput WITH label files: ConstructionLabelWriter.write_batch on c2 frames (64 x 1080p,
100 instances) to a tmpfs directory.  python tools/writer_files_bench.py [batches]"""
import json, shutil, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from constructionsceneposeestimation_b200 import synthetic
from constructionsceneposeestimation_b200.writer import ConstructionLabelWriter
batches = int(sys.argv[1]) if len(sys.argv) > 1 else 4
uniq = synthetic.make_batch(synthetic.CONFIGS["c2"], 8)
for fr in uniq:   # pinned host arrays, as a capture loop would hold them
    fr["instance_segmentation"]["data"] = torch.from_numpy(fr["instance_segmentation"]["data"].view("int32")).pin_memory().numpy()
for fmts, threads in ((("json",), 1), (("json",), None), (("json", "yolo"), None), (("yolo",), None), ((), None)):
    out = tempfile.mkdtemp(dir="/dev/shm")
    w = ConstructionLabelWriter(out, formats=fmts, split_people=True, io_threads=threads)
    def batch(k):
        frames = []
        for i in range(64):
            fr = dict(uniq[i % 8]); fr["frame_id"] = k * 64 + i; fr.pop("distance_to_image_plane", None); frames.append(fr)
        return frames
    w.write_batch(batch(0)); w.flush()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(1, batches + 1):
        w.write_batch(batch(k))
    w.flush(); dt = time.perf_counter() - t0
    w.on_final_frame()
    print(json.dumps({"formats": list(fmts), "io_threads": w.io_threads, "frames": batches * 64, "ms_per_frame": round(dt / (batches * 64) * 1e3, 3),
                      "frames_per_s": round(batches * 64 / dt, 1)}), flush=True)
    shutil.rmtree(out, ignore_errors=True)
