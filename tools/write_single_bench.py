"""Per-frame latency of the drop-in call: ConstructionLabelWriter.write(data) one 1080p frame at a time (the
reference's capture loop shape), label_%06d.json to tmpfs.  python tools/write_single_bench.py [frames]"""
import cProfile, json, pstats, shutil, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from constructionsceneposeestimation_b200 import synthetic
from constructionsceneposeestimation_b200.writer import ConstructionLabelWriter
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
uniq = synthetic.make_batch(synthetic.CONFIGS["c2"], 8)
for fr in uniq:
    fr["instance_segmentation"]["data"] = torch.from_numpy(fr["instance_segmentation"]["data"].view("int32")).pin_memory().numpy()
    fr.pop("distance_to_image_plane", None)
for fmts in (("json",), ()):
    out = tempfile.mkdtemp(dir="/dev/shm")
    w = ConstructionLabelWriter(out, formats=fmts, split_people=True)
    def frame(i):
        fr = dict(uniq[i % 8]); fr["frame_id"] = i; return fr
    for i in range(10):
        w.write(frame(i))
    w.flush(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10, 10 + n):
        w.write(frame(i))
    w.flush(); dt = time.perf_counter() - t0
    print(json.dumps({"formats": list(fmts), "frames": n, "ms_per_write": round(dt / n * 1e3, 3), "frames_per_s": round(n / dt, 1)}), flush=True)
    if fmts:
        pr = cProfile.Profile(); pr.enable()
        for i in range(10 + n, 10 + n + 100):
            w.write(frame(i))
        w.flush(); pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
    w.on_final_frame(); shutil.rmtree(out, ignore_errors=True)
