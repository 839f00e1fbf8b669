"""Host-side serialisation of emitted records: the reference's label JSON, COCO and YOLO.

SURVEY §8a rows R7 / S6.  ``reference_label`` keeps the reference's ``label_%06d.json`` schema
(gcd.py:2056-2064 frame dict, gcd.py:1938-1946 object dict, ``json.dump(indent=2,
ensure_ascii=False)`` gcd.py:613) key for key and only ADDS fields.
"""
from __future__ import annotations

import json
import math
import os
from typing import Union, Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np

from ._lib import OBJ_POSE_VALID
from .classes import CLASS_NAMES, CLASS_TABLE, SceneObject


def _finite_list(a) -> list:
    return [float(v) if math.isfinite(v) else None for v in np.asarray(a, dtype=np.float64).ravel()]


def object_entry(rec: np.void, obj: SceneObject, keypoints: Optional[Mapping] = None) -> Dict[str, object]:
    """One element of ``objects`` — reference keys first (gcd.py:1938-1946), extras after."""
    pose = rec["pose"]
    valid = bool(int(rec["flags"]) & OBJ_POSE_VALID)
    entry: Dict[str, object] = {
        "inst_idx": int(rec["inst_idx"]),
        "class_id": int(rec["class_id"]),
        "class_name": obj.class_name,
        "center": _finite_list(pose[7:10]),
        "size": _finite_list(pose[10:13]),
        "rotation": _finite_list(pose[13:16]) if valid else None,
        "prim_path": obj.prim_path,
        # ---- additions ([SPEC] stages) ----
        "pixel_count": int(rec["count"]),
        "bbox_2d_tight": [int(rec["x_min"]), int(rec["y_min"]), int(rec["x_max"]), int(rec["y_max"])],
        "bbox_2d_loose": [int(v) for v in rec["loose"]],
        "occlusion": float(rec["occlusion"]),
        "truncation": float(rec["truncation"]),
        "fill": float(rec["fill"]),
        "bbox_3d_projected": [[_f(u), _f(v)] for u, v in rec["uv"]],
        "bbox_3d_depth": _finite_list(rec["z"]),
        "pose_in_camera": {
            "translation": _finite_list(pose[0:3]),
            "quaternion_xyzw": _finite_list(pose[3:7]) if valid else None,
        },
        "flags": int(rec["flags"]),
    }
    if keypoints is not None:
        entry["keypoints"] = keypoints
    return entry


def _f(v) -> Optional[float]:
    v = float(v)
    return v if math.isfinite(v) else None


def reference_label(frame_id: int, camera_pose: Sequence[float], camera_params: Mapping, height: int, width: int,
                    records: np.ndarray, objects: Sequence[SceneObject],
                    keypoints_by_slot: Optional[Mapping[int, Mapping]] = None) -> Dict[str, object]:
    """The frame dict of gcd.py:2056-2064."""
    kps = keypoints_by_slot or {}
    objs = [object_entry(r, objects[int(r["inst_idx"])], kps.get(int(r["inst_idx"]))) for r in records]
    return {
        "frame_id": int(frame_id),
        "camera_pose": [float(v) for v in camera_pose],
        "camera_params": dict(camera_params),
        "objects": objs,
        "instance_mask_shape": [int(height), int(width)],
        "num_objects": len(objs),
        "class_mapping": dict(CLASS_TABLE),
    }


def dump_label_json(label: Mapping, path) -> None:
    """Same call as the reference's save_label_json (gcd.py:608-613)."""
    with open(path, "w", encoding="utf-8") as f:
        json.dump(label, f, indent=2, ensure_ascii=False)


def nested_json(value, level: int = 1) -> bytes:
    """``value`` as json.dump(indent=2, ensure_ascii=False) writes it ``level`` containers deep."""
    text = json.dumps(value, indent=2, ensure_ascii=False)
    return text.replace("\n", "\n" + "  " * level).encode("utf-8")


_CLASS_MAPPING_JSON = None


def slot_string_table(objects: Sequence[SceneObject]):
    """(blob, int32 offsets [2 * N + 1]): per slot the JSON literals of class_name and prim_path."""
    parts: List[bytes] = []
    offsets = [0]
    for o in objects:
        for text in (o.class_name, o.prim_path):
            parts.append(json.dumps(text, ensure_ascii=False).encode("utf-8"))
            offsets.append(offsets[-1] + len(parts[-1]))
    return b"".join(parts), np.asarray(offsets, dtype=np.int32)


def label_json_bytes(frame_id: int, camera_pose: Sequence[float], camera_params: Mapping, height: int, width: int,
                     records: np.ndarray, objects: Sequence[SceneObject], slot_strings=None,
                     keypoints: Optional[np.ndarray] = None, visibility: Optional[np.ndarray] = None,
                     person_slots: Optional[Sequence[int]] = None) -> bytes:
    """Native formatter (libcspe ``cspe_format_label_json_host``): the UTF-8 text
    ``json.dumps(reference_label(...), indent=2, ensure_ascii=False)`` gives, byte for byte, straight from
    the D2H record buffer.  ``slot_strings`` caches ``slot_string_table(objects)``; ``keypoints`` f64 [P,J,2] /
    ``visibility`` u8 [P,J] / ``person_slots`` (slot of person p, -1 = none) add the keypoint blocks."""
    global _CLASS_MAPPING_JSON
    from . import _lib

    lib = _lib.load()
    if _CLASS_MAPPING_JSON is None:
        _CLASS_MAPPING_JSON = nested_json(dict(CLASS_TABLE))
    blob, offsets = slot_strings if slot_strings is not None else slot_string_table(objects)
    records = np.ascontiguousarray(records)
    pose = np.ascontiguousarray(camera_pose, dtype=np.float64)
    if pose.shape != (7,):
        raise ValueError(f"camera_pose must have 7 entries, got shape {pose.shape}")
    n, num_slots = len(records), len(objects)
    kp_ptr = vis_ptr = pos_ptr = None
    P = J = 0
    if keypoints is not None and person_slots is not None:
        keypoints = np.ascontiguousarray(keypoints, dtype=np.float64)
        visibility = np.ascontiguousarray(visibility, dtype=np.uint8)
        P, J = visibility.shape
        person_of_slot = np.full(max(num_slots, 1), -1, dtype=np.int32)
        for p_idx in range(min(P, len(person_slots))):
            if 0 <= person_slots[p_idx] < num_slots:
                person_of_slot[person_slots[p_idx]] = p_idx
        kp_ptr, vis_ptr, pos_ptr = keypoints.ctypes.data, visibility.ctypes.data, person_of_slot.ctypes.data
    params_json = nested_json(dict(camera_params))
    # the bound the library checks: frame part + fragments + per record 4096 + its strings + 128 per joint
    cap = 2 * 4096 + len(_CLASS_MAPPING_JSON) + len(params_json) + n * (4096 + 128 * J) + int(offsets[-1])
    buf = np.empty(cap, dtype=np.uint8)
    rc = lib.cspe_format_label_json_host(records.ctypes.data if n else None, n, int(frame_id), pose.ctypes.data,
                                         params_json, _CLASS_MAPPING_JSON, blob,
                                         offsets.ctypes.data, num_slots, int(height), int(width), kp_ptr, vis_ptr,
                                         pos_ptr, P, J, buf.ctypes.data, cap)
    _lib.check("cspe_format_label_json_host", rc)
    return buf[:rc].tobytes()


def yolo_lines(records: np.ndarray) -> List[str]:
    """``class cx cy w h`` normalised to the image, one line per kept object."""
    return [
        f"{int(r['class_id'])} {r['yolo'][0]:.6f} {r['yolo'][1]:.6f} {r['yolo'][2]:.6f} {r['yolo'][3]:.6f}"
        for r in records
    ]


def yolo_text_batch(records: np.ndarray, n_out: np.ndarray, frames: Optional[int] = None):
    """Native formatter (libcspe ``cspe_format_yolo_host``): records RECORD_DTYPE [B,N] + n_out int32 [B]
    -> (bytes buffer, int64 offsets [frames+1]); frame f's YOLO text is buf[offsets[f]:offsets[f+1]],
    byte-identical to ``"\n".join(yolo_lines(...)) + "\n"``."""
    from . import _lib

    lib = _lib.load()
    records = np.ascontiguousarray(records)
    n_out = np.ascontiguousarray(n_out, dtype=np.int32)
    B, N = records.shape
    frames = B if frames is None else frames
    cap = int(n_out[:frames].sum()) * 96 + 16
    buf = np.empty(cap, dtype=np.uint8)
    offsets = np.empty(frames + 1, dtype=np.int64)
    rc = lib.cspe_format_yolo_host(records.ctypes.data, n_out.ctypes.data, B, N, frames, buf.ctypes.data, cap,
                                   offsets.ctypes.data)
    _lib.check("cspe_format_yolo_host", rc)
    return buf[:rc], offsets


def write_files(directory: str, prefix: str, suffix: str, first_id: int, data: np.ndarray, sizes: np.ndarray,
                count: Optional[int] = None, digits: int = 6) -> int:
    """Native batch file writer (libcspe ``cspe_write_files_host``, no GIL): file j =
    ``<directory>/<prefix><first_id + j:0{digits}d><suffix>`` holds ``data[j, :sizes[j]]`` (data u8 [B, stride],
    e.g. the D2H buffer of ``cspe_format_yolo``).  Returns the bytes written."""
    from . import _lib

    lib = _lib.load()
    if data.ndim != 2 or data.dtype != np.uint8 or not data.flags.c_contiguous:
        raise ValueError("data must be a C-contiguous uint8 [B, stride] array")
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    count = len(sizes) if count is None else int(count)
    if count > data.shape[0] or count > len(sizes):
        raise ValueError("count exceeds the rows of data / sizes")
    rc = lib.cspe_write_files_host(os.fsencode(directory), prefix.encode(), int(digits), suffix.encode(), int(first_id),
                                   count, data.ctypes.data, data.shape[1], sizes.ctypes.data)
    _lib.check("cspe_write_files_host", rc)
    return int(rc)


def concat_rows(data: np.ndarray, sizes: np.ndarray, count: Optional[int] = None, as_array: bool = False):
    """``data[j, :sizes[j]]`` for j < count, back to back (libcspe ``cspe_concat_rows_host``): the per-frame label text
    of a D2H buffer (``cspe_format_yolo`` / ``cspe_format_coco``) as one chunk."""
    from . import _lib

    lib = _lib.load()
    if data.ndim != 2 or data.dtype != np.uint8 or not data.flags.c_contiguous:
        raise ValueError("data must be a C-contiguous uint8 [B, stride] array")
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    count = len(sizes) if count is None else int(count)
    if count > data.shape[0] or count > len(sizes):
        raise ValueError("count exceeds the rows of data / sizes")
    cap = int(np.clip(sizes[:count], 0, None).sum())
    out = np.empty(cap, dtype=np.uint8)
    rc = lib.cspe_concat_rows_host(data.ctypes.data, data.shape[1], sizes.ctypes.data, count, out.ctypes.data, cap)
    _lib.check("cspe_concat_rows_host", rc)
    return out[:rc] if as_array else out[:rc].tobytes()   # as_array: no second copy (a uint8 array is bytes-like)


def coco_categories() -> List[Dict[str, object]]:
    return [{"id": i, "name": n, "supercategory": "construction"} for i, n in enumerate(CLASS_NAMES)]


def coco_annotations(records: np.ndarray, image_id: int, first_ann_id: int,
                     keypoints_by_slot: Optional[Mapping[int, Mapping]] = None) -> List[Dict[str, object]]:
    """COCO: bbox = [x_min, y_min, w, h] of the tight box, area = pixel count, iscrowd = 0;
    keypoints = [x, y, v] * J with num_keypoints = sum(v > 0)."""
    kps = keypoints_by_slot or {}
    anns = []
    for i, r in enumerate(records):
        x0, y0, x1, y1 = int(r["x_min"]), int(r["y_min"]), int(r["x_max"]), int(r["y_max"])
        ann: Dict[str, object] = {
            "id": first_ann_id + i,
            "image_id": int(image_id),
            "category_id": int(r["class_id"]),
            "bbox": [x0, y0, x1 - x0 + 1, y1 - y0 + 1] if int(r["count"]) > 0 else [0, 0, 0, 0],
            "area": int(r["count"]),
            "iscrowd": 0,
            # ratios to six decimals: what a label consumer needs, and printable on the device (csrc/repr6.h)
            "occlusion": round(float(r["occlusion"]), 6),
            "truncation": round(float(r["truncation"]), 6),
        }
        kp = kps.get(int(r["inst_idx"]))
        if kp is not None:
            ann["keypoints"] = kp["keypoints"]
            ann["num_keypoints"] = kp["num_keypoints"]
        anns.append(ann)
    return anns


def coco_annotations_text(records: np.ndarray, n_out: np.ndarray, image_ids: Sequence[int], first_ann_id: int,
                          keypoints: Optional[np.ndarray] = None, visibility: Optional[np.ndarray] = None,
                          person_slots: Optional[Sequence[Sequence[int]]] = None):
    """Native formatter (libcspe ``cspe_format_coco_host``): the annotations of a whole batch as the text
    ``json.dumps([...])`` puts between its brackets — byte-identical to dumping ``coco_annotations`` frame by
    frame.  records RECORD_DTYPE [B,N], n_out int32 [B], image_ids per frame; keypoints f64 [B,P,J,2] /
    visibility u8 [B,P,J] / person_slots[f][p] = slot of person p (or -1).  Returns (bytes, number of annotations)."""
    from . import _lib

    lib = _lib.load()
    records = np.ascontiguousarray(records)
    n_out = np.ascontiguousarray(n_out, dtype=np.int32)
    B, N = records.shape
    frames = len(image_ids)
    ids = np.ascontiguousarray(image_ids, dtype=np.int64)
    kp_ptr = vis_ptr = pos_ptr = None
    P = J = 0
    if keypoints is not None and person_slots is not None:
        keypoints = np.ascontiguousarray(keypoints, dtype=np.float64)
        visibility = np.ascontiguousarray(visibility, dtype=np.uint8)
        _, P, J = visibility.shape
        person_of_slot = np.full((B, max(N, 1)), -1, dtype=np.int32)
        for f in range(min(frames, len(person_slots))):
            for p_idx, slot in enumerate(person_slots[f][:P]):
                if 0 <= slot < N:
                    person_of_slot[f, slot] = p_idx
        kp_ptr, vis_ptr, pos_ptr = keypoints.ctypes.data, visibility.ctypes.data, person_of_slot.ctypes.data
    total = int(n_out[:frames].sum())
    cap = 64 + total * (512 + 128 * J)
    buf = np.empty(cap, dtype=np.uint8)
    rc = lib.cspe_format_coco_host(records.ctypes.data, n_out.ctypes.data, B, N, frames, ids.ctypes.data,
                                   int(first_ann_id), kp_ptr, vis_ptr, pos_ptr, P, J, buf.ctypes.data, cap)
    _lib.check("cspe_format_coco_host", rc)
    return buf[:rc].tobytes(), total


def coco_images_text(image_ids: Sequence[int], width: int, height: int) -> bytes:
    """The ``images`` entries of a run of frames as the text ``json.dumps([coco_image(...), ...])`` puts between
    its brackets (a 100 k-frame sweep would otherwise build and dump 100 k dicts).  A contiguous ``range`` goes
    through the native formatter (libcspe ``cspe_format_coco_images_host``), anything else through the Python
    statement of the same text."""
    w, h = int(width), int(height)
    if isinstance(image_ids, range) and image_ids.step == 1 and len(image_ids) > 0:
        from . import _lib

        lib = _lib.load()
        cap = 160 * len(image_ids)
        buf = np.empty(cap, dtype=np.uint8)
        rc = lib.cspe_format_coco_images_host(int(image_ids.start), len(image_ids), w, h, buf.ctypes.data, cap)
        _lib.check("cspe_format_coco_images_host", rc)
        return buf[:rc].tobytes()
    return ", ".join(f'{{"id": {int(i)}, "width": {w}, "height": {h}, "file_name": "rgb_{int(i):06d}.png"}}'
                     for i in image_ids).encode("ascii")


def write_coco_file(path, images: Sequence[Union[Mapping, bytes]], annotation_chunks: Sequence[bytes],
                    joined: bool = False) -> None:
    """The bytes json.dump({"images": [...], "annotations": [...], "categories": [...]}, f) writes, with the
    annotations given as natively formatted chunks (``coco_annotations_text``) and the images either as dicts
    (``coco_image``) or as pre-formatted chunks (``coco_images_text``)."""
    with open(path, "wb") as f:
        if images and isinstance(images[0], (bytes, bytearray)):
            f.write(b'{"images": [' + b", ".join(c for c in images if len(c)) + b'], "annotations": [')
        else:
            f.write(b'{"images": ' + json.dumps(list(images)).encode("ascii") + b', "annotations": [')
        # joined=True: the chunks already carry their ", " separators (device formatter)
        # (chunks are bytes-like: bytes or uint8 arrays straight from the D2H buffer — len(), not truthiness)
        sep = b"" if joined else b", "
        first = True
        for c in annotation_chunks:
            if len(c) == 0:
                continue
            if not first and sep:
                f.write(sep)
            f.write(c)
            first = False
        f.write(b'], "categories": ' + json.dumps(coco_categories()).encode("ascii") + b"}")


def coco_keypoint_block(kp: np.ndarray, vis: np.ndarray) -> Dict[str, object]:
    """kp f64 [J,2], vis u8 [J] -> {"keypoints": [x,y,v]*J, "num_keypoints": n} (x,y zeroed when v == 0)."""
    flat: List[float] = []
    for (x, y), v in zip(kp, vis):
        v = int(v)
        if v == 0 or not (math.isfinite(x) and math.isfinite(y)):
            flat += [0.0, 0.0, 0]
        else:
            flat += [float(x), float(y), v]
    return {"keypoints": flat, "num_keypoints": int((np.asarray(vis) > 0).sum())}


def coco_image(image_id: int, width: int, height: int, file_name: str) -> Dict[str, object]:
    return {"id": int(image_id), "width": int(width), "height": int(height), "file_name": file_name}


def pointcloud_annotator_rows(pcd) -> Optional[np.ndarray]:
    """The (N, 6) ``x y z r g b`` matrix the reference writes for a Replicator ``pointcloud`` annotator payload
    (``save_pointcloud_with_rgb``, gcd.py:715-769 — the capture loop's FIRST choice for ``pointcloud_%06d.txt``, before
    the depth-map fallback of gcd.py:1729-1759), or None where the reference writes nothing.

    ``pcd["data"]`` is xyz (N, 3) — a single point may arrive flat (3,); ``pcd["pointRgb"]`` is (N, 3 | 4) colours (the
    alpha column is dropped), a single flat colour, or absent / empty / malformed -> white (255); when the two lengths
    disagree both are cut to the shorter one.  The matrix is ``np.hstack([xyz, rgb])`` exactly as the reference builds
    it (dtype promotion included), so ``np.savetxt(fmt='%.6f')`` — or ``cspe_format_fixed6`` on its float64 copy — prints
    the reference's bytes."""
    try:
        if pcd is None or "data" not in pcd:
            return None
        xyz = pcd["data"]
        if xyz is None or len(xyz) == 0:
            return None
        xyz = np.asarray(xyz)
        if xyz.ndim == 1:
            if len(xyz) != 3:
                return None
            xyz = xyz.reshape(1, 3)
        n = xyz.shape[0]
        white = lambda: np.ones((n, 3)) * 255      # float64, like the reference's default
        colours = pcd.get("pointRgb") if hasattr(pcd, "get") else None
        if colours is None:
            rgb = white()
        else:
            colours = np.asarray(colours)
            if colours.ndim == 1:
                rgb = colours[:3].reshape(1, 3) if len(colours) in (3, 4) else white()
            else:
                rgb = colours[:, :3] if colours.shape[1] >= 3 else colours
        if xyz.shape[0] != rgb.shape[0]:
            m = min(xyz.shape[0], rgb.shape[0])
            xyz, rgb = xyz[:m], rgb[:m]
        return np.hstack([xyz, rgb])
    except Exception:   # the reference wraps the whole function in try / except and skips the file (gcd.py:771-776)
        return None
