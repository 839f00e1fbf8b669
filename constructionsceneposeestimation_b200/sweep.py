"""Frame-range sweep driver (BASELINE config 5): stream N frames per rank through the pipeline,
emit YOLO / COCO labels and all-gather the per-class histogram at the end.

    torchrun --nproc-per-node 8 -m constructionsceneposeestimation_b200.sweep --frames 100000 --emit yolo --out DIR

Each rank owns the contiguous global frame range ``sharding.frame_range(rank, world, frames)``
and cycles a device-resident pool of synthetic annotator frames (>= 64 x 1080p = 531 MB, far
beyond L2) so every batch streams from HBM.  Records come back through pinned host buffers,
double-buffered: batch k+1 runs on the GPU while the host formats batch k.
"""
from __future__ import annotations

import argparse
import json
import os
import time
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, classes, formats, sharding, synthetic
from .camera import pack_camera
from .pipeline import LabelPipeline


def build_host_tables(frames, split_people=True):
    res = classes.ObjectRootResolver(split_people=split_people)
    per = []
    for fr in frames:
        paths = fr["bounding_box_3d"]["info"]["primPaths"]
        objs = classes.aggregate_objects(paths, res)
        per.append((objs, classes.record_index_for(objs, paths),
                    classes.id_to_slot(fr["instance_segmentation"]["info"]["idToLabels"], objs, res)))
    B = len(frames)
    N = max(1, max(len(p[0]) for p in per))
    R = max(1, max(len(fr["bounding_box_3d"]["data"]) for fr in frames))
    L = max(1, max((max(p[2]) if p[2] else 0) for p in per) + 1)
    L = (L + 3) & ~3  # 16-byte LUT rows (K1 stages each frame's LUT with a bulk copy)
    lut = np.full((B, L), -1, dtype=np.int32)
    obj_record = np.full((B, N), -1, dtype=np.int32)
    slot_class = np.full((B, N), -1, dtype=np.int32)
    records = np.zeros((B, R, _lib.BBOX3D_RECORD_BYTES), dtype=np.uint8)
    cam = np.zeros((B, _lib.CAM_STRIDE))
    for i, (fr, (objs, rec_idx, mapping)) in enumerate(zip(frames, per)):
        for k, v in mapping.items():
            lut[i, k] = v
        obj_record[i, : len(objs)] = rec_idx
        slot_class[i, : len(objs)] = [o.class_id for o in objs]
        r = np.ascontiguousarray(fr["bounding_box_3d"]["data"])
        records[i, : len(r)] = r.view(np.uint8).reshape(len(r), -1)
        pack_camera(fr["camera_pose"], fr["camera_params"], out=cam[i])
    return lut, obj_record, slot_class, records, cam, [p[0] for p in per]


def run_sweep(num_frames: int, rank: int = 0, world: int = 1, device: Optional[torch.device] = None,
              pool_frames: int = 64, config: str = "c2", emit: Optional[str] = "yolo", out_dir: Optional[str] = None,
              use_graph: bool = True) -> Dict[str, object]:
    """Annotate this rank's share of ``num_frames``; returns counters, timings and the histogram."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    lo, hi = sharding.frame_range(rank, world, num_frames)
    spec = synthetic.CONFIGS[config]
    pool = synthetic.make_batch(spec, pool_frames, first_frame=rank * pool_frames)
    lut, obj_record, slot_class, records, cam, objects = build_host_tables(pool)
    H, W = pool[0]["instance_segmentation"]["data"].shape
    B, N = pool_frames, obj_record.shape[1]
    with torch.cuda.device(device):
        pipe = LabelPipeline(B, H, W, N, records.shape[1], lut.shape[1], device, use_graph=use_graph)
        pipe.mask.copy_(torch.from_numpy(np.stack([f["instance_segmentation"]["data"] for f in pool]).view(np.int32)))
        pipe.lut.copy_(torch.from_numpy(lut))
        pipe.obj_record.copy_(torch.from_numpy(obj_record))
        pipe.slot_class.copy_(torch.from_numpy(slot_class))
        pipe.records_in.copy_(torch.from_numpy(records))
        pipe.cam.copy_(torch.from_numpy(cam))
        host = [(torch.empty(pipe.records.shape, dtype=torch.uint8, pin_memory=True),
                 torch.empty((B,), dtype=torch.int32, pin_memory=True), torch.cuda.Event()) for _ in range(2)]
        torch.cuda.synchronize(device)

        label_dir = None
        if out_dir is not None and emit is not None:
            label_dir = os.path.join(out_dir, "labels")
            os.makedirs(label_dir, exist_ok=True)
        coco_imgs: List[dict] = []
        coco_anns: List[bytes] = []
        coco_count = 0
        slot_strings = [formats.slot_string_table(o) for o in objects] if emit == "json" else None
        io_pool = None
        if emit == "json":
            from concurrent.futures import ThreadPoolExecutor

            io_pool = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1), thread_name_prefix="cspe-io")
        emitted = 0
        hist_host = np.zeros(_lib.NUM_CLASSES, dtype=np.int64)
        batches = sharding.batches(lo, hi, B)
        t0 = time.perf_counter()

        def launch(k: int) -> None:
            rec_h, n_h, ev = host[k % 2]
            pipe.run()
            rec_h.copy_(pipe.records, non_blocking=True)
            n_h.copy_(pipe.n_out, non_blocking=True)
            ev.record()

        def consume(k: int) -> None:
            nonlocal emitted, hist_host, coco_count
            rec_h, n_h, ev = host[k % 2]
            ev.synchronize()
            s, e = batches[k]
            recs_all = rec_h.numpy().view(_lib.RECORD_DTYPE).reshape(B, N)
            n_all = n_h.numpy()
            nf = e - s
            valid = np.arange(N)[None, :] < n_all[:nf, None]
            emitted += int(n_all[:nf].sum())
            hist_host += np.bincount(recs_all["class_id"][:nf][valid], minlength=_lib.NUM_CLASSES)[: _lib.NUM_CLASSES]
            if emit == "yolo":
                buf, off = formats.yolo_text_batch(recs_all, n_all, nf)   # native formatter (libcspe, f3)
                if label_dir is not None:
                    raw = buf.tobytes()
                    for j in range(nf):
                        with open(os.path.join(label_dir, f"label_{s + j:06d}.txt"), "wb") as fh:
                            fh.write(raw[off[j]:off[j + 1]])
                return
            if emit == "json":   # native formatter (libcspe, f3) on the I/O threads: label_%06d.json, gcd.py:2071
                def one(j: int) -> int:
                    text = formats.label_json_bytes(s + j, pool[j]["camera_pose"], pool[j]["camera_params"], H, W,
                                                    recs_all[j, : n_all[j]], objects[j], slot_strings[j])
                    if label_dir is not None:
                        with open(os.path.join(label_dir, f"label_{s + j:06d}.json"), "wb") as fh:
                            fh.write(text)
                    return len(text)

                sum(io_pool.map(one, range(nf)))
                return
            if emit == "coco":   # native formatter, one call per batch
                ids = list(range(s, e))               # global frame ids; pipeline frames are batch-relative
                coco_imgs.extend(formats.coco_image(fid, W, H, f"rgb_{fid:06d}.png") for fid in ids)
                text, count = formats.coco_annotations_text(recs_all, n_all, ids, coco_count + 1)
                coco_anns.append(text)
                coco_count += count

        # The device histogram (K4) counts every frame of every launched batch; a trailing partial
        # batch still runs the whole pool, so the frames this rank OWNS are counted on the host from
        # the consumed records and cross-checked against the device when all batches were full.
        pipe.class_hist.zero_()
        for k in range(len(batches)):
            launch(k)
            if k > 0:
                consume(k - 1)
        if batches:
            consume(len(batches) - 1)
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        if io_pool is not None:
            io_pool.shutdown()
        if batches and all(e - s == B for s, e in batches):
            dev_hist = pipe.class_hist.cpu().numpy()
            if not np.array_equal(dev_hist, hist_host):
                raise RuntimeError(f"class histogram mismatch: device {dev_hist.tolist()} vs host {hist_host.tolist()}")
        hist_dev = torch.from_numpy(hist_host).to(device)

        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            gathered = sharding.all_gather_histogram(hist_dev)
        else:
            gathered = hist_host.reshape(1, -1)
        if emit == "coco" and out_dir is not None:
            formats.write_coco_file(os.path.join(out_dir, f"coco_rank{rank:02d}.json"), coco_imgs, coco_anns)
    return {"rank": rank, "world": world, "frames": hi - lo, "frame_range": [lo, hi], "records": emitted,
            "seconds": dt, "frames_per_s": (hi - lo) / dt if dt > 0 else 0.0, "class_hist_rank": hist_host.tolist(),
            "class_hist_total": gathered.sum(axis=0).tolist(), "class_hist_per_rank": gathered.tolist()}


def main() -> int:
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=100_000)
    ap.add_argument("--pool", type=int, default=64)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--emit", default="yolo", choices=["yolo", "coco", "json", "none"])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run_sweep(args.frames, rank, world, torch.device("cuda", local), args.pool, args.config,
                    None if args.emit == "none" else args.emit, args.out)
    if world > 1:
        t = torch.tensor([res["seconds"]], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["seconds_max_over_ranks"] = float(t.item())
        res["frames_per_s_all_ranks"] = args.frames / float(t.item())
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res), flush=True)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
