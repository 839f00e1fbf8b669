"""Replicator-Writer-style front end of the annotation hot path.

``ConstructionLabelWriter.write(data)`` takes the annotator dict a Replicator writer receives
(the same objects the reference's capture loop pulls by hand: depth gcd.py:1681, bbox3d
gcd.py:1780-1790 / 1916-1922, instance segmentation gcd.py:1818-1842, camera gcd.py:1599 and
2036-2045) and produces what the reference writes at gcd.py:2055-2072 — ``label_%06d.json``
(same schema, extra fields added) and ``instance_mask_%06d.npy`` — plus COCO / YOLO records.

Host Python only prepares small tables (prim path -> slot / class / record, camera block);
pixels are touched exclusively by the CUDA kernels in libcspe.so.  No CPU fallback.

Error conventions follow the reference: a bad object is skipped or flagged, never fatal
(gcd.py:2024-2027); ``None`` / empty annotators are tolerated (gcd.py:1682, 1788, 1919).
"""
from __future__ import annotations

import json
import os
import warnings
import weakref
from collections.abc import Mapping   # isinstance() against typing.Mapping goes through a pure-Python __instancecheck__
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, formats, ops
from .quality import FrameQualityLog, depth_quality_from_stats
from ._lib import BBOX3D_DTYPE, CAM_STRIDE, NUM_CLASSES, RECORD_DTYPE
from .camera import (DEFAULT_FAR, DEFAULT_NEAR, camera_params as default_camera_params, from_replicator_camera_params,
                     is_replicator_camera_params, pack_camera, pack_cameras)
from .classes import (CLASS_NAMES, RECORD_APPROX_BIT, ObjectRootResolver, SceneObject, aggregate_objects, id_to_slot,
                      label_path, pack_union, record_index_for, union_members)

ArrayLike = Union[np.ndarray, torch.Tensor]

_IDENTITY_POSE = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]


@dataclass
class FrameTables:
    """Host tables of one frame (cached per distinct scene signature)."""
    objects: List[SceneObject]
    obj_record: np.ndarray      # int32 [N]
    slot_class: np.ndarray      # int32 [N]
    lut_ids: np.ndarray         # int64 [K] instance ids that map to a slot
    lut_slots: np.ndarray       # int32 [K]
    max_id: int
    slot_strings: Optional[Tuple[bytes, np.ndarray]] = None   # JSON literals of class_name / prim_path per slot (lazy)
    # record_fallback="union": [(slot, [mesh record indices])] of the objects that get an object-level record
    union: Sequence[Tuple[int, Sequence[int]]] = ()
    n_paths: int = 0            # len(primPaths) the table was built from (union indices are relative to it)


@dataclass
class _HostBatch:
    """Host side of one batch (``ConstructionLabelWriter._host_tables``)."""
    tables: List[FrameTables]
    frame_ids: List[int]
    poses: List[Sequence[float]]
    params_list: List[Mapping]
    block: "_TableBlock"          # [lut | obj_record | slot_class | cam | records | union tables], pinned
    same_tables: bool              # one scene for the whole batch: the LUT / union tables have ONE row
    N: int                         # slots per frame
    R0: int                        # bbox3d records per frame (union records follow at R0 .. R0 + U)
    U: int                         # object-level records built on the device per frame (record_fallback="union")
    frame_base: int
    contiguous_ids: bool


@dataclass
class BatchLabels:
    """Result of one batch: device outputs, pinned host copies and the event that fences them."""
    frame_ids: List[int]
    tables: List[FrameTables]
    height: int
    width: int
    camera_poses: List[Sequence[float]]
    camera_params: List[Mapping]
    _rec_host: torch.Tensor
    _nout_host: torch.Tensor
    _event: torch.cuda.Event
    _kp_host: Optional[torch.Tensor] = None
    _vis_host: Optional[torch.Tensor] = None
    person_slots: Optional[List[List[int]]] = None
    device_outputs: Dict[str, torch.Tensor] = field(default_factory=dict)
    _depth_stats_host: Optional[torch.Tensor] = None
    _depth_viz_host: Optional[torch.Tensor] = None
    _synced: bool = False
    rgb_images: Optional[List[Optional[ArrayLike]]] = None   # per-frame RGB(A) for the point-cloud file
    _yolo: Optional[Tuple[np.ndarray, np.ndarray]] = None
    # recorded on the writer's stream after the last kernel / copy that reads the caller's annotator buffers
    inputs_consumed: Optional[torch.cuda.Event] = None
    missing_masks: Optional[List[int]] = None   # batch indices whose instance_segmentation annotator was absent
    # (N, 6) x y z r g b rows of the frames that brought a ``pointcloud`` annotator payload (gcd.py:1720-1727), else None
    pointcloud_rows: Optional[List[Optional[np.ndarray]]] = None

    def wait_inputs_consumed(self) -> "BatchLabels":
        """Block the host until the GPU no longer reads the annotator buffers this batch was built from (pinned
        host arrays being copied, device tensors used in place).  Producers that run on the CUDA stream that was
        current when the batch was submitted need not call this: that stream already waits for the event."""
        if self.inputs_consumed is not None:
            self.inputs_consumed.synchronize()
        return self

    def synchronize(self) -> "BatchLabels":
        if not self._synced:
            self._event.synchronize()
            self._synced = True
        return self

    @property
    def n_out(self) -> np.ndarray:
        self.synchronize()
        return self._nout_host.numpy()

    def records(self, f: int) -> np.ndarray:
        """Structured records (``_lib.RECORD_DTYPE``) of batch frame ``f``, in inst_idx order."""
        self.synchronize()
        n = int(self._nout_host[f])
        return self._rec_host[f].numpy().view(RECORD_DTYPE).reshape(-1)[:n]

    def keypoints(self, f: int) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        if self._kp_host is None:
            return None
        self.synchronize()
        return self._kp_host[f].numpy(), self._vis_host[f].numpy()

    def keypoints_by_slot(self, f: int) -> Dict[int, Dict[str, object]]:
        kv = self.keypoints(f)
        if kv is None or self.person_slots is None:
            return {}
        kp, vis = kv
        return {slot: formats.coco_keypoint_block(kp[p], vis[p])
                for p, slot in enumerate(self.person_slots[f]) if slot >= 0 and p < kp.shape[0]}

    def depth_quality(self, f: int) -> Optional[Dict[str, object]]:
        """The dict the reference's DataQualityLogger.log_depth stores per frame (gcd.py:333-342)."""
        if self._depth_stats_host is None:
            return None
        self.synchronize()
        return depth_quality_from_stats(self._depth_stats_host[f].numpy().view(_lib.DEPTH_STATS_DTYPE)[0])

    def depth_image(self, f: int) -> Optional[np.ndarray]:
        """JET visualisation of the depth map, uint8 BGR [H,W,3] (gcd.py:1691-1709); needs "depth_png" in formats."""
        if self._depth_viz_host is None:
            return None
        self.synchronize()
        return self._depth_viz_host[f].numpy()

    def reference_label(self, f: int) -> Dict[str, object]:
        return formats.reference_label(self.frame_ids[f], self.camera_poses[f], self.camera_params[f], self.height,
                                       self.width, self.records(f), self.tables[f].objects,
                                       self.keypoints_by_slot(f))

    def yolo_text(self, f: int) -> memoryview:
        """``class cx cy w h`` lines of frame ``f`` (native formatter, one call per batch, cached)."""
        if self._yolo is None:
            self.synchronize()
            B, N = self._rec_host.shape[0], self._rec_host.shape[1]
            recs = self._rec_host.numpy().view(RECORD_DTYPE).reshape(B, N)
            self._yolo = formats.yolo_text_batch(recs, self._nout_host.numpy())
        buf, offsets = self._yolo
        return memoryview(buf)[int(offsets[f]): int(offsets[f + 1])]

    def coco_text(self, first_annotation_id: int) -> Tuple[bytes, int]:
        """COCO annotation objects of the whole batch, ", "-joined (native formatter); returns (text, count)."""
        self.synchronize()
        B, N = self._rec_host.shape[0], self._rec_host.shape[1]
        recs = self._rec_host.numpy().view(RECORD_DTYPE).reshape(B, N)
        kp = vis = None
        if self._kp_host is not None and self.person_slots is not None:
            kp, vis = self._kp_host.numpy(), self._vis_host.numpy()
        return formats.coco_annotations_text(recs, self._nout_host.numpy(), self.frame_ids, first_annotation_id, kp, vis,
                                             self.person_slots if kp is not None else None)

    def label_json(self, f: int) -> bytes:
        """The text of ``label_%06d.json`` (= json.dumps(self.reference_label(f), indent=2, ensure_ascii=False)),
        formatted natively from the D2H record buffer."""
        t = self.tables[f]
        if t.slot_strings is None:
            t.slot_strings = formats.slot_string_table(t.objects)
        kv = self.keypoints(f) if self.person_slots is not None else None
        kp, vis = kv if kv is not None else (None, None)
        return formats.label_json_bytes(self.frame_ids[f], self.camera_poses[f], self.camera_params[f], self.height,
                                        self.width, self.records(f), t.objects, t.slot_strings, kp, vis,
                                        self.person_slots[f] if kv is not None else None)

    def __len__(self) -> int:
        return len(self.frame_ids)


class _FrameList(list):
    """Per-frame dicts cut from a stacked batch dict; remembers the stacked pixel arrays so that they reach the
    device in one copy (or none, when they already live there)."""
    stacked: Dict[str, ArrayLike]
    canonical: bool = False


def _unstack(data: Mapping) -> _FrameList:
    """``{"instance_segmentation": {"data": [B,H,W], "info": ...}, ...}`` -> list of B frame dicts (views, no copy).
    Every annotator may carry one ``info`` shared by the batch or a list of B; ``camera_pose`` is [B,7] (or one
    pose), ``camera_params`` one dict or a list, ``frame_id`` a list or the id of the first frame."""
    seg = data.get("instance_segmentation")
    masks = _payload(seg)
    if masks is None or masks.ndim != 3:
        raise ValueError("stacked batch: instance_segmentation data must be [B,H,W]")
    B = int(masks.shape[0])

    def per_frame(value, i):
        if isinstance(value, (list, tuple)) and len(value) == B:
            return value[i]
        return value

    rows_of: Dict[int, Sequence] = {}

    def rows(payload):
        """payload[i] for every i: one unbind for a torch tensor (B Python-level indexing calls on a CUDA tensor cost
        more than everything else in this function together)."""
        r = rows_of.get(id(payload))
        if r is None:
            r = payload.unbind(0) if isinstance(payload, torch.Tensor) and payload.shape[0] == B else payload
            rows_of[id(payload)] = r
        return r

    def cut(annot, i):
        if annot is None:
            return None
        if type(annot) is dict or isinstance(annot, Mapping):
            payload = annot.get("data")
            out = {"data": None if payload is None else rows(payload)[i]}
            if "info" in annot:
                out["info"] = per_frame(annot["info"], i)
            return out
        return rows(annot)[i]

    frames = _FrameList()
    fid = data.get("frame_id")
    pose = data.get("camera_pose")
    if pose is not None:
        pose = np.asarray(pose, dtype=np.float64)
        pose_rows = pose.tolist() if pose.ndim == 2 else None
    fid_is_seq = fid is not None and np.ndim(fid) > 0
    for i in range(B):
        fr: Dict[str, object] = {}
        for name in ("instance_segmentation", "distance_to_image_plane", "bounding_box_3d", "rgb"):
            if data.get(name) is not None:
                fr[name] = cut(data[name], i)
        sk = data.get("skeleton_data")
        if sk is not None:
            j = sk.get("globalTranslations") if isinstance(sk, Mapping) else sk
            fr["skeleton_data"] = {"globalTranslations": j[i]}
        if pose is not None:
            fr["camera_pose"] = pose_rows[i] if pose_rows is not None else list(pose)
        if data.get("camera_params") is not None:
            fr["camera_params"] = per_frame(data["camera_params"], i)
        if fid is not None:
            fr["frame_id"] = int(fid[i]) if fid_is_seq else int(fid) + i
        frames.append(fr)
    frames.stacked = {"instance_segmentation": masks}
    frames.canonical = True   # keys are plain annotator names already
    depth = _payload(data.get("distance_to_image_plane"))
    if depth is not None and getattr(depth, "ndim", 0) == 3:
        frames.stacked["distance_to_image_plane"] = depth
    return frames


def tables_cache_key(prim_paths: Sequence[str], id_to_labels: Mapping) -> Tuple:
    """Hashable signature of a scene: the prim paths and (id, labelled path) pairs.  Replicator's idToLabels
    values are strings or ``{"class": ...}`` dicts and its keys may be strings (gcd.py:1826-1837).  The common case
    (every value a string) is keyed on the raw key / value tuples — no per-entry Python work; dict values are not
    hashable and take the normalising path."""
    try:
        key = (tuple(prim_paths), tuple(id_to_labels.keys()), tuple(id_to_labels.values()))
        hash(key)
        return key
    except TypeError:
        return tuple(prim_paths), tuple((str(k), label_path(v)) for k, v in id_to_labels.items())


def _info(annot) -> Mapping:
    if type(annot) is dict or isinstance(annot, Mapping):
        info = annot.get("info")
        return info if (type(info) is dict or isinstance(info, Mapping)) else {}
    return {}


def _payload(annot):
    if type(annot) is dict or isinstance(annot, Mapping):
        return annot.get("data")
    return annot


_ANNOTATOR_KEYS = ("instance_segmentation", "distance_to_image_plane", "bounding_box_3d", "camera_params",
                   "skeleton_data", "rgb", "camera_pose", "pointcloud")


_NAME_CACHE: Dict[str, Tuple[Optional[str], Optional[str]]] = {}


def _annotator_name(key: str) -> Tuple[Optional[str], Optional[str]]:
    """(canonical annotator name, render-product suffix) of a Replicator payload key: writers attached to a
    render product receive ``"<annotator>-<render product>"`` keys, and the ``*_fast`` variants deliver the same
    payload as their plain annotator.  (None, None) for keys that are not annotators of this path."""
    hit = _NAME_CACHE.get(key)
    if hit is not None:
        return hit
    base, _, suffix = key.partition("-")
    if base.endswith("_fast"):
        base = base[:-5]
    if base == "LdrColor":
        base = "rgb"
    out = (base, suffix or None) if base in _ANNOTATOR_KEYS else (None, None)
    if len(_NAME_CACHE) < 4096:
        _NAME_CACHE[key] = out
    return out


def split_render_products(data: Mapping) -> List[Dict[str, object]]:
    """One frame dict per render product of a Replicator ``write(data)`` payload, annotator keys canonicalised.
    A multi-camera rig (BASELINE config 4) arrives as ONE payload with suffixed keys; its cameras are returned in
    sorted render-product order (the order ``sharding.rig_frame_range`` flattens (rig frame, camera) pairs in)."""
    shared: Dict[str, object] = {}
    per: Dict[str, Dict[str, object]] = {}
    for key, value in data.items():
        name, suffix = _annotator_name(key) if isinstance(key, str) else (None, None)
        if name is None:
            shared[key] = value
        elif suffix is None:
            shared[name] = value
        else:
            per.setdefault(suffix, {})[name] = value
    if not per:
        return [shared]
    frames = []
    for suffix in sorted(per):
        fr = dict(shared)
        fr.update(per[suffix])
        fr["render_product"] = suffix
        frames.append(fr)
    return frames


def skeleton_joints(sk) -> Optional[np.ndarray]:
    """World joint positions float32 [P,J,3] from a ``skeleton_data`` payload, or None.  Accepted shapes: an array
    [P,J,3] (or [J,3] for one skeleton); ``{"globalTranslations": array}`` (also under ``"data"``); a list of
    per-skeleton dicts each holding ``globalTranslations`` [J,3]; a dict of parallel per-skeleton lists
    (``{"skeletonData" | "skeletons": [...]}``); or any of these as a JSON string.  (Replicator's exact payload
    cannot be checked offline — see INTEGRATION.md; whatever does not parse is treated as "no skeletons".)"""
    if sk is None:
        return None
    if isinstance(sk, (str, bytes)):
        try:
            sk = json.loads(sk)
        except (ValueError, TypeError):
            return None
    if isinstance(sk, Mapping):
        for key in ("globalTranslations", "global_translations"):
            if sk.get(key) is not None:
                return skeleton_joints(sk[key])
        for key in ("data", "skeletonData", "skeletons", "skeleton_data"):
            if sk.get(key) is not None:
                return skeleton_joints(sk[key])
        return None
    if isinstance(sk, (list, tuple)) and sk and isinstance(sk[0], (Mapping, str, bytes)):
        per = [skeleton_joints(item) for item in sk]
        per = [p[0] if p is not None and p.ndim == 3 and p.shape[0] == 1 else p for p in per]
        if any(p is None or p.ndim != 2 for p in per) or len({p.shape for p in per}) != 1:
            return None
        return np.stack(per).astype(np.float32, copy=False)
    try:
        j = np.asarray(sk, dtype=np.float32)
    except (ValueError, TypeError):
        return None
    if j.ndim == 2 and j.shape[-1] == 3:
        j = j[None]
    return j if j.ndim == 3 and j.shape[-1] == 3 else None


class _TableBlock:
    """The small per-batch host tables laid out in ONE pinned block — [lut | obj_record | slot_class | cam | records]
    — so that they reach the device in a single H2D copy into a matching device block."""

    def __init__(self, writer: "ConstructionLabelWriter", lut_rows: int, L: int, B: int, N: int, R: int,
                 union_shape: Tuple[int, int, int] = (0, 0, 0)):
        u_rows, U1, M = union_shape   # record_fallback="union": offsets int32 [u_rows][U1], members int32 [u_rows][M]
        sizes = [lut_rows * L * 4, B * N * 4, B * N * 4, B * CAM_STRIDE * 8, B * R * BBOX3D_DTYPE.itemsize,
                 u_rows * U1 * 4, u_rows * M * 4]
        offs = [0]
        for sz in sizes:
            offs.append((offs[-1] + sz + 15) & ~15)
        self.buffers: List[Tuple[Tuple, torch.Tensor]] = []
        self.host = writer._take_pinned("tables", (offs[-1],), torch.uint8, self.buffers)
        self.device = writer.device
        h = self.host.numpy()
        self.lut = h[offs[0]: offs[0] + sizes[0]].view(np.int32).reshape(lut_rows, L)
        self.obj_record = h[offs[1]: offs[1] + sizes[1]].view(np.int32).reshape(B, N)
        self.slot_class = h[offs[2]: offs[2] + sizes[2]].view(np.int32).reshape(B, N)
        self.cam = h[offs[3]: offs[3] + sizes[3]].view(np.float64).reshape(B, CAM_STRIDE)
        self.records = h[offs[4]: offs[4] + sizes[4]].reshape(B, R, BBOX3D_DTYPE.itemsize)
        self.union_offsets = h[offs[5]: offs[5] + sizes[5]].view(np.int32).reshape(u_rows, U1)
        self.union_members = h[offs[6]: offs[6] + sizes[6]].view(np.int32).reshape(u_rows, M)
        self._offs, self._sizes, self._shapes, self._union_shape = offs, sizes, (lut_rows, L, B, N, R), union_shape

    def upload(self):
        lut_rows, L, B, N, R = self._shapes
        d = self.host.to(self.device, non_blocking=True)
        self.device_block = d
        o, z = self._offs, self._sizes
        return (d[o[0]: o[0] + z[0]].view(torch.int32).view(lut_rows, L),
                d[o[1]: o[1] + z[1]].view(torch.int32).view(B, N),
                d[o[2]: o[2] + z[2]].view(torch.int32).view(B, N),
                d[o[4]: o[4] + z[4]].view(B, R, BBOX3D_DTYPE.itemsize),
                d[o[3]: o[3] + z[3]].view(torch.float64).view(B, CAM_STRIDE))

    def union_on_device(self, d: torch.Tensor):
        """(offsets, members) views of the device block ``d`` (the storage ``upload`` returned views of)."""
        u_rows, U1, M = self._union_shape
        o, z = self._offs, self._sizes
        return (d[o[5]: o[5] + z[5]].view(torch.int32).view(u_rows, U1),
                d[o[6]: o[6] + z[6]].view(torch.int32).view(u_rows, M))


class ConstructionLabelWriter:
    """Drop-in writer for the reference's per-frame label path.

    Parameters mirror the knobs the reference hard-codes: output directory layout
    (``labels/`` as gcd.py:40), clipping range (gcd.py:1437), and the run-time crane part map
    (gcd.py:124).  ``formats`` selects what ``write`` serialises: ``"json"`` (reference schema),
    ``"yolo"``, ``"coco"``, ``"mask"`` (the real instance mask instead of the reference's -1
    placeholder, gcd.py:2066-2069), ``"depth_png"`` (JET depth image, gcd.py:1691-1709), ``"depth_csv"``
    (``depth/depth_%06d.csv``, the np.savetxt text of gcd.py:1688) and ``"pointcloud"``
    (``pointcloud/pointcloud_%06d.txt`` from depth + ``data["rgb"]``, gcd.py:1729-1759) — both texts are
    formatted on the GPU (``cspe_format_fixed6``) and only their bytes cross PCIe — and ``"rgb_png"``
    (``rgb/rgb_%06d.png`` from ``data["rgb"]``, gcd.py:1669-1674; the RGB(A) -> BGR conversion runs on the GPU,
    PNG encoding is cv2's as in the reference).  When a
    frame carries ``distance_to_image_plane`` the depth-quality statistics of the reference's
    logger (gcd.py:314-359) are computed on the GPU and returned by ``BatchLabels.depth_quality``.
    """

    annotators = ["instance_segmentation", "distance_to_image_plane", "bounding_box_3d", "camera_params",
                  "skeleton_data"]

    def __init__(self, output_dir: Optional[str] = None, device: Union[str, torch.device, None] = None,
                 formats: Sequence[str] = ("json",), min_pixels: int = 1, keypoint_tolerance: float = 0.15,
                 near: float = DEFAULT_NEAR, far: float = DEFAULT_FAR, split_people: bool = False,
                 record_fallback: str = "first_mesh", crane_part_map: Optional[Mapping] = None,
                 rank: int = 0, world_size: int = 1, max_pending: int = 2, quality_log: bool = True,
                 io_threads: Optional[int] = None):
        _lib.load()  # fail loudly right here if the CUDA library is missing
        if not torch.cuda.is_available():
            raise _lib.CspeLibraryError("ConstructionLabelWriter needs a CUDA device (no CPU fallback exists)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.output_dir = output_dir
        self.formats = tuple(formats)
        self.min_pixels = int(min_pixels)
        self.keypoint_tolerance = float(keypoint_tolerance)
        self.near, self.far = float(near), float(far)
        if record_fallback not in ("first_mesh", "reference", "union"):
            raise ValueError(f"unknown record fallback {record_fallback!r}")
        self.record_fallback = record_fallback
        self.resolver = ObjectRootResolver(crane_part_map, split_people=split_people)
        self.rank, self.world_size = rank, world_size
        self.max_pending = max_pending
        self.max_lut_entries = 1 << 28     # 1 GiB of id -> slot table per batch
        # per-frame files of a batch are formatted and written by a small thread pool
        self.io_threads = min(16, os.cpu_count() or 1) if io_threads is None else max(1, int(io_threads))
        self._io_pool = None
        self._tables_cache: Dict[Tuple, FrameTables] = {}
        self._free_pinned: Dict[Tuple, List[torch.Tensor]] = {}
        self._warned_empty_lut = False
        self._next_frame_id = 0
        self._next_rig_frame = 0
        self._pending: List[Tuple[BatchLabels, Optional[List[np.ndarray]]]] = []
        self._coco_images: List[Dict] = []
        self._coco_annotation_text: List[bytes] = []   # natively formatted, one chunk per batch
        self._coco_annotation_count = 0
        self.frames_written = 0
        self.objects_total = 0
        self.depth_quality_log: List[Dict[str, object]] = []
        # the reference's DataQualityLogger (gcd.py:236-464): logs/generation_summary.json
        self.quality: Optional[FrameQualityLog] = None
        if quality_log:
            self.quality = FrameQualityLog(os.path.join(output_dir, "logs") if output_dir is not None else None)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream(device=self.device)
            self.class_hist = torch.zeros((NUM_CLASSES,), dtype=torch.int64, device=self.device)
        if output_dir is not None:
            os.makedirs(os.path.join(output_dir, "labels"), exist_ok=True)

    # ------------------------------------------------------------------ host tables
    def frame_tables(self, prim_paths: Sequence[str], id_to_labels: Mapping) -> FrameTables:
        key = tables_cache_key(prim_paths, id_to_labels)
        hit = self._tables_cache.get(key)
        if hit is not None:
            return hit
        objects = aggregate_objects(prim_paths, self.resolver)
        rec_idx = np.asarray(record_index_for(objects, prim_paths, self.record_fallback), dtype=np.int32).reshape(-1)
        slot_class = np.asarray([o.class_id for o in objects], dtype=np.int32).reshape(-1)
        mapping = id_to_slot(id_to_labels, objects, self.resolver)
        ids = np.fromiter(mapping.keys(), dtype=np.int64, count=len(mapping))
        slots = np.fromiter(mapping.values(), dtype=np.int32, count=len(mapping))
        ok = (ids >= 0) & (ids < (1 << 32))
        ids, slots = ids[ok], slots[ok]
        tables = FrameTables(objects, rec_idx, slot_class, ids, slots, int(ids.max()) if ids.size else -1)
        tables.n_paths = len(prim_paths)
        if self.record_fallback == "union":
            tables.union = union_members(objects, prim_paths)
        if len(self._tables_cache) > 4096:
            self._tables_cache.clear()
        self._tables_cache[key] = tables
        return tables

    # ------------------------------------------------------------------ public surface
    def write(self, data: Mapping) -> None:
        """One Replicator payload (``Writer.write`` signature): one frame, or — when the payload carries several
        render products (``"<annotator>-<render product>"`` keys) — one frame per camera of the rig, numbered
        ``frame_id * cameras + camera`` in sorted render-product order."""
        frames = split_render_products(data)
        if len(frames) > 1:
            fid = data.get("frame_id")
            base = (int(fid) if fid is not None else self._next_rig_frame) * len(frames)
            self._next_rig_frame = (int(fid) if fid is not None else self._next_rig_frame) + 1
            for c, fr in enumerate(frames):
                fr["frame_id"] = base + c
        self.write_batch(frames)

    def write_batch(self, frames: Union[Sequence[Mapping], Mapping]) -> BatchLabels:
        """Annotate B frames in one set of launches and queue them for serialisation.  ``frames`` is a list of
        frame dicts (what ``write`` takes) or ONE dict of stacked annotators (``[B,H,W]`` masks / depth, ``[B,R]``
        records, ``[B,7]`` poses; see ``_unstack``) — stacked pixel arrays go to the device in a single copy."""
        if isinstance(frames, Mapping):
            frames = _unstack(frames)
        labels = self.annotate_batch(frames)
        # Serialisation is deferred by up to max_pending batches, so whatever it needs besides the records is
        # snapshotted NOW into the writer's own device memory (the mask batch inside annotate_batch, RGB here):
        # the caller may reuse its annotator buffers as soon as inputs_consumed has fired.
        if {"pointcloud", "rgb_png"} & set(self.formats) and self.output_dir is not None:
            labels.rgb_images = self._snapshot_rgb(frames, labels)
        if "pointcloud" in self.formats and self.output_dir is not None:
            # the capture loop's first choice for pointcloud_%06d.txt is the pointcloud annotator itself
            # (gcd.py:1720-1727); its small host arrays are turned into the (N, 6) matrix right here
            rows = [self._pointcloud_payload_rows(fr) for fr in frames]
            if any(r is not None for r in rows):
                labels.pointcloud_rows = rows
        self._pending.append((labels, None))
        while len(self._pending) > self.max_pending:
            self._serialise(*self._pending.pop(0))
        return labels

    @staticmethod
    def _pointcloud_payload_rows(fr: Mapping) -> Optional[np.ndarray]:
        """formats.pointcloud_annotator_rows of the frame's ``pointcloud`` annotator payload, if it brought one.  The
        reference reads ``pointRgb`` next to ``data`` (gcd.py:735); Replicator versions that deliver it under ``info``
        are accepted too.  CUDA tensors are brought to the host (the cloud is a few thousand points)."""
        payload = None
        for key, value in fr.items():   # keys may still carry a render-product suffix here
            if isinstance(key, str) and _annotator_name(key)[0] == "pointcloud":
                payload = value
        if not isinstance(payload, Mapping) or payload.get("data") is None:
            return None
        host = lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
        colours = payload.get("pointRgb")
        if colours is None and isinstance(payload.get("info"), Mapping):
            colours = payload["info"].get("pointRgb")
        pcd = {"data": host(payload["data"])}
        if colours is not None:
            pcd["pointRgb"] = host(colours)
        return formats.pointcloud_annotator_rows(pcd)

    def _snapshot_rgb(self, frames, labels: BatchLabels) -> List[Optional[torch.Tensor]]:
        dev = self.device
        caller_stream = torch.cuda.current_stream(dev)
        out: List[Optional[torch.Tensor]] = []
        with torch.cuda.device(dev), torch.cuda.stream(self.stream):
            for fr in frames:
                rgb = None
                for key, value in fr.items():   # keys may still carry a render-product suffix here
                    if isinstance(key, str) and _annotator_name(key)[0] == "rgb":
                        rgb = _payload(value)
                if rgb is None:
                    out.append(None)
                elif isinstance(rgb, torch.Tensor) and rgb.is_cuda:
                    rgb.record_stream(self.stream)
                    out.append(rgb.to(dev).clone())
                else:
                    src = rgb if isinstance(rgb, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rgb))
                    out.append(src.to(dev, non_blocking=True))
            consumed = torch.cuda.Event()
            consumed.record(self.stream)
        caller_stream.wait_event(consumed)
        labels.inputs_consumed = consumed
        return out

    def flush(self) -> None:
        while self._pending:
            self._serialise(*self._pending.pop(0))

    def on_final_frame(self) -> Dict[str, object]:
        """Flush pending frames, gather the per-class histogram across ranks and write the summary."""
        self.flush()
        if self._io_pool is not None:
            self._io_pool.shutdown()
            self._io_pool = None
        hist = self.gather_class_histogram()
        summary = {
            "frames": self.frames_written,
            # the reference logger's object counter (gcd.py:361-372, 392-394)
            "object_count": {"total": self.objects_total,
                             "per_frame_avg": self.objects_total / self.frames_written if self.frames_written else 0},
            "depth_quality": self.depth_quality_log,
            "rank": self.rank,
            "world_size": self.world_size,
            "class_histogram": {CLASS_NAMES[i]: int(hist["total"][i]) for i in range(NUM_CLASSES)},
            "class_histogram_per_rank": hist["per_rank"].tolist(),
        }
        if self.output_dir is not None:
            if "coco" in self.formats:
                formats.write_coco_file(os.path.join(self.output_dir, f"coco_rank{self.rank:02d}.json"),
                                        self._coco_images, self._coco_annotation_text)
            if self.rank == 0:
                with open(os.path.join(self.output_dir, "label_summary.json"), "w", encoding="utf-8") as f:
                    json.dump(summary, f, indent=2)
            if self.quality is not None:   # gcd.py:2090; one file per rank beyond rank 0
                name = "generation_summary.json" if self.world_size == 1 else f"generation_summary_rank{self.rank:02d}.json"
                os.makedirs(self.quality.log_dir, exist_ok=True)
                self.quality.save_summary(os.path.join(self.quality.log_dir, name))
        if self.quality is not None:
            summary["quality"] = self.quality.summary()["statistics"]
        return summary

    def gather_class_histogram(self) -> Dict[str, np.ndarray]:
        """S7: all-gather of the int64[10] histogram (NCCL when torch.distributed is up)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .sharding import all_gather_histogram

            per_rank = all_gather_histogram(self.class_hist)
        else:
            per_rank = self.class_hist.detach().cpu().numpy().reshape(1, NUM_CLASSES)
        return {"per_rank": per_rank, "total": per_rank.sum(axis=0)}

    # ------------------------------------------------------------------ host tables of a batch
    def _host_tables(self, frames: Sequence[Mapping], H: int, W: int, missing: Sequence[int]) -> "_HostBatch":
        """Everything the kernels need besides the pixels, for B normalised frame dicts: per-frame scene tables
        (cached), bbox3d records, camera blocks, frame ids — laid out in ONE pinned block (``_TableBlock``).  Pure host
        work (numpy); ``tests/test_host_logic.py`` runs it on the CPU against the test helpers' statement of the same
        tables."""
        B = len(frames)
        tables: List[FrameTables] = []
        rec_arrays: List[Optional[np.ndarray]] = []
        cams = np.zeros((B, CAM_STRIDE), dtype=np.float64)
        frame_ids: List[int] = []
        poses, params_list, clips = [], [], []
        # info dicts that are the SAME OBJECT for several frames of this call (one info shared by a stacked batch,
        # or a pool of frames repeated) are resolved once: {(id(primPaths), id(idToLabels), records): tables}
        seen: Dict[Tuple[int, int, int], FrameTables] = {}
        for i, fr in enumerate(frames):
            seg = fr.get("instance_segmentation")
            bbox = fr.get("bounding_box_3d")
            raw_paths = _info(bbox).get("primPaths", None) or ()   # tolerated empty, gcd.py:1788
            recs = _payload(bbox)
            id_to_labels = _info(seg).get("idToLabels", None) or {}
            n_paths = len(raw_paths)
            n_recs = n_paths if recs is None else len(recs)
            ident = (id(raw_paths), id(id_to_labels), min(n_recs, n_paths))
            hit = seen.get(ident) if n_paths else None
            if hit is not None:
                tables.append(hit)
                if n_recs > n_paths:
                    recs = recs[:n_paths]
            else:
                prim_paths = list(raw_paths)
                if recs is not None and n_recs != n_paths:
                    n = min(n_recs, n_paths)
                    recs, prim_paths = recs[:n], prim_paths[:n]
                tables.append(self.frame_tables(prim_paths, id_to_labels))
                if n_paths:
                    seen[ident] = tables[-1]
                prim_paths_nonempty = bool(prim_paths)
            if hit is None and prim_paths_nonempty and not tables[-1].lut_ids.size and i not in missing \
                    and not self._warned_empty_lut:
                self._warned_empty_lut = True
                warnings.warn("bounding_box_3d lists prims but no idToLabels entry of instance_segmentation resolves to "
                              "one of them: every object of such frames has 0 pixels and is dropped (min_pixels >= 1)")
            rec_arrays.append(recs if recs is not None and len(recs) else None)
            params = fr.get("camera_params") or default_camera_params(W, H)
            if "width" not in params or "height" not in params:
                params = {**params, "width": params.get("width", W), "height": params.get("height", H)}
            pose = fr.get("camera_pose")
            if pose is None:
                pose = _IDENTITY_POSE
            clips.append(fr.get("_clip") or (self.near, self.far))
            poses.append(pose)
            params_list.append(params)
            fid = fr.get("frame_id")
            frame_ids.append(int(fid) if fid is not None else self._next_frame_id + i)
        pack_cameras(poses, params_list, clips, out=cams)
        frame_base = frame_ids[0]
        contiguous_ids = all(frame_ids[i] == frame_base + i for i in range(B))
        self._next_frame_id = max(self._next_frame_id, max(frame_ids) + 1)

        N = max(1, max(len(t.objects) for t in tables))
        R = max(1, max((len(r) for r in rec_arrays if r is not None), default=1))
        same_tables = all(t is tables[0] for t in tables)
        L = max(1, max(t.max_id for t in tables) + 1)
        L = (L + 3) & ~3  # 16-byte LUT rows
        lut_rows = 1 if same_tables else B
        if L * lut_rows > self.max_lut_entries:
            # the id -> slot table is dense (4 bytes per possible id, per distinct scene in the batch); instance ids
            # are prim indices in Replicator, so this only trips on ids that are not instance ids
            raise ValueError(f"instance ids up to {L - 1} need a {L * lut_rows * 4 / 2**20:.0f} MiB id->slot table "
                             f"(limit {self.max_lut_entries * 4 / 2**20:.0f} MiB, max_lut_entries)")
        # record_fallback="union": U object-level records per frame are built on the device behind the frame's own R
        U = max(len(t.union) for t in tables)
        R0 = R
        if U:
            u_off, u_mem, _ = pack_union([tables[0].union] if same_tables else [t.union for t in tables])
            R = R0 + U
        # small tables go up as ONE pinned block: [lut | obj_record | slot_class | records | cam | union tables]
        blk = _TableBlock(self, lut_rows, L, B, N, R, (u_off.shape[0], u_off.shape[1], u_mem.shape[1]) if U else (0, 0, 0))
        lut, obj_record, slot_class, rec_bytes = blk.lut, blk.obj_record, blk.slot_class, blk.records
        lut.fill(-1)
        obj_record.fill(-1)
        slot_class.fill(-1)
        blk.cam[:] = cams
        if same_tables:   # one scene for the whole batch: broadcast its rows
            t = tables[0]
            n = len(t.objects)
            obj_record[:, :n] = t.obj_record
            slot_class[:, :n] = t.slot_class
            lut[0, t.lut_ids] = t.lut_slots
        else:   # frames that share a scene (a pool of frames repeated, a static scene with a few variants) fill together
            rows_of: Dict[int, List[int]] = {}
            for i, t in enumerate(tables):
                rows_of.setdefault(id(t), []).append(i)
            for rows in rows_of.values():
                t = tables[rows[0]]
                n = len(t.objects)
                if len(rows) == 1:
                    i = rows[0]
                    obj_record[i, :n] = t.obj_record
                    slot_class[i, :n] = t.slot_class
                    lut[i, t.lut_ids] = t.lut_slots
                else:
                    r = np.asarray(rows)
                    obj_record[r, :n] = t.obj_record
                    slot_class[r, :n] = t.slot_class
                    lut[r[:, None], t.lut_ids[None, :]] = t.lut_slots
        if U:   # union objects point behind the frame's own records (their table entry is relative to len(primPaths))
            blk.union_offsets[:] = u_off
            blk.union_members[:] = u_mem
            for i, t in ([(slice(None), tables[0])] if same_tables else enumerate(tables)):
                for u, (slot, _) in enumerate(t.union):
                    obj_record[i, slot] = (R0 + u) | RECORD_APPROX_BIT
        item = BBOX3D_DTYPE.itemsize
        dt0 = rec_arrays[0].dtype if type(rec_arrays[0]) is np.ndarray else None
        if dt0 is not None and dt0.itemsize == item and \
                all(type(r) is np.ndarray and r.ndim == 1 and len(r) == R and r.dtype == dt0 and r.flags.c_contiguous
                    for r in rec_arrays):
            # every frame brings R records: one C-level gather of BYTES straight into the pinned block (concatenating the
            # structured arrays themselves goes through numpy's per-field copy: 1.2 ms instead of 0.17 ms for 64 x 222)
            np.concatenate([r.view(np.uint8) for r in rec_arrays], out=rec_bytes.reshape(-1))
        else:
            rec_bytes.fill(0)
            for i, r in enumerate(rec_arrays):
                if r is not None:
                    r = np.ascontiguousarray(r)
                    if r.dtype.itemsize != item:
                        raise ValueError(f"frame {i}: bounding_box_3d record itemsize {r.dtype.itemsize} != 96")
                    rec_bytes[i, : len(r)] = r.view(np.uint8).reshape(len(r), -1)
                else:
                    obj_record[i, :] = -1

        return _HostBatch(tables, frame_ids, poses, params_list, blk, same_tables, N, R0, U, frame_base, contiguous_ids)

    # ------------------------------------------------------------------ the hot path
    def annotate_batch(self, frames: Union[Sequence[Mapping], Mapping]) -> BatchLabels:
        """Enqueue the whole path for B frames (list of frame dicts or one stacked dict, like ``write_batch``)
        and return at once; the result is fenced by ``BatchLabels.synchronize()``.

        Buffer contract (INTEGRATION.md "Input lifetime"): the annotator arrays are read by the GPU after this
        call returns — pinned host arrays by asynchronous H2D copies, CUDA tensors in place.  The CUDA stream that
        is current on entry is made to wait for ``BatchLabels.inputs_consumed``, so producers that write those
        buffers from that stream are ordered automatically; anything else (a host thread refilling a pinned staging
        buffer, another stream) must call ``BatchLabels.wait_inputs_consumed()`` first."""
        stacked: Dict[str, ArrayLike] = {}
        if isinstance(frames, Mapping):
            frames = _unstack(frames)
            stacked = frames.stacked
        if len(frames) == 0:
            raise ValueError("annotate_batch needs at least one frame")
        if getattr(frames, "canonical", False):   # cut from a stacked dict: only the camera payload may need work
            frames = [self._normalise_camera(fr) for fr in frames]
        else:
            frames = [self._normalise(fr) for fr in frames]
        B = len(frames)
        dev = self.device

        # ---- the big copy first: enqueue the H2D of the masks (PCIe-bound, ~10 ms for 64 x 1080p)
        # before building the host tables, so the table work hides under it -----------------
        masks: List[Optional[ArrayLike]] = [_payload(fr.get("instance_segmentation")) for fr in frames]
        missing = [i for i, m in enumerate(masks) if m is None or getattr(m, "ndim", 0) != 2]
        shape = next(((int(m.shape[0]), int(m.shape[1])) for i, m in enumerate(masks) if i not in missing), None)
        if shape is None:   # no mask at all: take the image size from depth, the camera, or the script default
            shape = self._fallback_shape(frames[0])
        H, W = shape
        for i, m in enumerate(masks):
            if i not in missing and tuple(m.shape) != (H, W):
                raise ValueError(f"frame {i}: mask shape {tuple(m.shape)} differs from {(H, W)} (one [H,W] resolution per batch)")
        caller_stream = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            # device-resident annotators (device="cuda" in Replicator) were produced on the caller's stream
            self.stream.wait_stream(caller_stream)
            with torch.cuda.stream(self.stream):
                if missing:   # an absent annotator gives an empty-label frame, never an error (gcd.py:1682,1788,1919)
                    d_mask = torch.zeros((B, H, W), dtype=torch.int32, device=dev)
                    present = [i for i in range(B) if i not in missing]
                    if present:
                        d_mask[present] = self._stack_to_device([masks[i] for i in present], torch.int32)
                    mask_owned = True
                else:
                    d_mask, mask_owned = self._stack_to_device(masks, torch.int32, stacked.get("instance_segmentation"),
                                                               return_owned=True)

        # ---- host: per-frame tables (numpy only; overlaps the mask copy enqueued above) ------------------
        hb = self._host_tables(frames, H, W, missing)
        tables, frame_ids, poses, params_list, blk = hb.tables, hb.frame_ids, hb.poses, hb.params_list, hb.block
        same_tables, N, R0, U, frame_base, contiguous_ids = hb.same_tables, hb.N, hb.R0, hb.U, hb.frame_base, hb.contiguous_ids

        # ---- device: uploads + kernels on the writer's stream ----------------------------
        owned: List[Tuple[Tuple, torch.Tensor]] = list(blk.buffers)
        with torch.cuda.device(dev), torch.cuda.stream(self.stream):
            d_lut, d_obj_record, d_slot_class, d_rec, d_cam = blk.upload()
            if same_tables:
                d_lut = d_lut[0]

            if U:
                d_uoff, d_umem = blk.union_on_device(blk.device_block)
                ops.union_records(d_rec, R0, d_uoff[0] if same_tables else d_uoff, d_umem[0] if same_tables else d_umem)
            scan = ops.mask_scan(d_mask, d_lut, N)
            uv, z, pose, loose, flags = ops.project_objects(d_rec, d_obj_record, d_cam)
            kp_host = vis_host = None
            person_slots = None
            d_kp = d_vis = d_depth = None
            joints_list = [skeleton_joints(fr.get("skeleton_data")) for fr in frames]
            depth_list = [_payload(fr.get("distance_to_image_plane")) for fr in frames]
            if all(j is not None and j.shape[0] > 0 for j in joints_list) and all(d is not None for d in depth_list) \
                    and len({tuple(j.shape) for j in joints_list}) == 1:
                d_depth = self._stack_to_device(depth_list, torch.float32, stacked.get("distance_to_image_plane"))
                d_joints = self._stack_to_device(joints_list, torch.float32)
                d_kp, _kz, d_vis = ops.keypoints(d_joints, d_depth, d_cam, self.keypoint_tolerance)
                person_slots = [self._person_slots(t, joints_list[i].shape[0]) for i, t in enumerate(tables)]
            stats_host = viz_host = None
            if all(d is not None for d in depth_list):
                if d_depth is None:
                    d_depth = self._stack_to_device(depth_list, torch.float32, stacked.get("distance_to_image_plane"))
                d_stats = ops.depth_stats(d_depth)                                  # f2, gcd.py:314-359
                stats_host = self._take_pinned("stats", tuple(d_stats.shape), torch.uint8, owned)
                stats_host.copy_(d_stats, non_blocking=True)
                if "depth_png" in self.formats:
                    d_viz = ops.depth_colormap(d_depth, self._jet_lut(), d_stats)   # f4, gcd.py:1691-1709
                    viz_host = self._take_pinned("viz", tuple(d_viz.shape), torch.uint8, owned)
                    viz_host.copy_(d_viz, non_blocking=True)
            rec_dev, n_out, _ = ops.emit(scan, uv, z, pose, loose, flags, d_slot_class, H, W, self.min_pixels,
                                         frame_base if contiguous_ids else 0, class_hist=self.class_hist)
            # everything that reads the caller's buffers has been enqueued: mark it, make the caller's stream wait
            # for it, and keep the allocator from recycling caller tensors we read in place
            keep_depth = d_depth is not None and bool({"depth_csv", "pointcloud"} & set(self.formats))
            keep_mask = "mask" in self.formats and self.output_dir is not None
            if keep_mask and not mask_owned:
                d_mask = d_mask.clone()      # deferred serialisation must not depend on the caller's buffer
            if keep_depth and self._is_caller_tensor(d_depth, depth_list, stacked.get("distance_to_image_plane")):
                d_depth = d_depth.clone()
            consumed = torch.cuda.Event()
            consumed.record(self.stream)
            # (per-frame views cut from a stacked CUDA tensor share its storage: recording the stacked tensor covers them)
            st_mask, st_depth = stacked.get("instance_segmentation"), stacked.get("distance_to_image_plane")
            for t in [st_mask, st_depth, *(masks if st_mask is None else ()), *(depth_list if st_depth is None else ())]:
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(self.stream)
            rec_host = self._take_pinned("rec", tuple(rec_dev.shape), torch.uint8, owned)
            nout_host = self._take_pinned("nout", tuple(n_out.shape), torch.int32, owned)
            rec_host.copy_(rec_dev, non_blocking=True)
            nout_host.copy_(n_out, non_blocking=True)
            if d_kp is not None:
                kp_host = self._take_pinned("kp", tuple(d_kp.shape), torch.float64, owned)
                vis_host = self._take_pinned("vis", tuple(d_vis.shape), torch.uint8, owned)
                kp_host.copy_(d_kp, non_blocking=True)
                vis_host.copy_(d_vis, non_blocking=True)
            event = torch.cuda.Event()
            event.record(self.stream)
        caller_stream.wait_event(consumed)

        labels = BatchLabels(frame_ids, tables, H, W, poses, params_list, rec_host, nout_host, event, kp_host,
                             vis_host, person_slots,
                             {"scan": scan, "uv": uv, "z": z, "pose": pose, "loose": loose, "flags": flags,
                              "records": rec_dev, "n_out": n_out, "cam": d_cam,
                              **({"depth": d_depth} if keep_depth else {}),
                              **({"mask": d_mask} if keep_mask else {})}, stats_host, viz_host)
        labels.inputs_consumed = consumed
        labels.missing_masks = missing
        # pinned blocks go back to the writer's free lists when the result object dies
        weakref.finalize(labels, self._give_back, owned, event)
        if not contiguous_ids:
            labels.synchronize()
            for f in range(B):  # frame field was written relative to 0
                labels._rec_host[f].numpy().view(RECORD_DTYPE).reshape(-1)["frame"][: int(nout_host[f])] = frame_ids[f]
        return labels

    # ------------------------------------------------------------------ input normalisation
    def _normalise(self, fr: Mapping) -> Dict[str, object]:
        """Canonical frame dict: annotator keys without render-product / ``_fast`` suffixes; Replicator's native
        ``camera_params`` payload turned into the reference's five-field dict plus the pose it implies (a separate
        ``camera_pose`` entry, as the reference passes it at gcd.py:1599, wins)."""
        out: Dict[str, object] = {}
        for key, value in fr.items():
            name, _ = _annotator_name(key) if isinstance(key, str) else (None, None)
            if name is None:
                out[key] = value
            elif name not in out or out[name] is None:
                out[name] = value
        return self._normalise_camera(out)

    @staticmethod
    def _normalise_camera(out: Dict[str, object]) -> Dict[str, object]:
        cp = out.get("camera_params")
        if cp is None:
            return out
        if isinstance(cp, Mapping) and "data" in cp and isinstance(cp["data"], Mapping):
            cp = cp["data"]
        if is_replicator_camera_params(cp):
            seg = _payload(out.get("instance_segmentation"))
            hw = (int(seg.shape[-1]), int(seg.shape[-2])) if seg is not None and getattr(seg, "ndim", 0) >= 2 else (None, None)
            pose7, params, clip = from_replicator_camera_params(cp, *hw)
            out["camera_params"] = params
            if out.get("camera_pose") is None:
                out["camera_pose"] = pose7
            if clip is not None:
                out["_clip"] = clip
        elif cp is not None:
            out["camera_params"] = cp
        return out

    @staticmethod
    def _fallback_shape(fr: Mapping) -> Tuple[int, int]:
        depth = _payload(fr.get("distance_to_image_plane"))
        if depth is not None and getattr(depth, "ndim", 0) == 2:
            return int(depth.shape[0]), int(depth.shape[1])
        params = fr.get("camera_params") or {}
        if params.get("width") and params.get("height"):
            return int(params["height"]), int(params["width"])
        return 720, 1280   # the script's capture resolution, gcd.py:46-47

    @staticmethod
    def _is_caller_tensor(t: torch.Tensor, parts: Sequence, stacked) -> bool:
        """True when ``t`` aliases a CUDA tensor the caller handed in (used in place, not copied)."""
        cands = [stacked] + list(parts)
        return any(isinstance(c, torch.Tensor) and c.is_cuda and c.untyped_storage().data_ptr() == t.untyped_storage().data_ptr()
                   for c in cands)

    # ------------------------------------------------------------------ buffer reuse
    def _take_pinned(self, kind: str, shape: Tuple[int, ...], dtype: torch.dtype, owned: List) -> torch.Tensor:
        """A pinned host tensor from the writer's free lists (allocated on first use); registered in ``owned`` so it
        returns to the list when the batch's result object is garbage collected."""
        key = (kind, shape, dtype)
        free = self._free_pinned.get(key) or []
        t = None
        for i, (cand, ev) in enumerate(free):
            if ev is None or ev.query():   # the batch that used it last has run to its end on the GPU
                t = cand
                del free[i]
                break
        if t is None:
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
        owned.append((key, t))
        return t

    def _give_back(self, owned: List, event: Optional[torch.cuda.Event] = None) -> None:
        for key, t in owned:
            free = self._free_pinned.setdefault(key, [])
            if len(free) < 4:
                free.append((t, event))

    # ------------------------------------------------------------------ helpers
    def _stack_to_device(self, arrays: Sequence[ArrayLike], dtype: torch.dtype,
                         stacked: Optional[ArrayLike] = None, return_owned: bool = False):
        """[B, ...] device tensor from per-frame host arrays / device tensors (async copies); ``stacked`` is the
        same data as one [B, ...] array when the caller has it (one copy, or none if it is on the device).
        ``return_owned=True`` also tells whether the result is the writer's own memory (False = the caller's CUDA
        tensor used in place)."""
        out, owned = self._stack_impl(arrays, dtype, stacked)
        return (out, owned) if return_owned else out

    def _stack_impl(self, arrays, dtype, stacked):
        if stacked is not None:
            if isinstance(stacked, torch.Tensor):
                t = stacked.view(torch.int32) if dtype == torch.int32 and stacked.dtype == torch.uint32 else stacked
            else:
                a = np.ascontiguousarray(stacked)
                t = torch.from_numpy(a.view(np.int32) if dtype == torch.int32 and a.dtype == np.uint32 else a)
            if t.dtype == dtype:
                in_place = t.is_cuda and t.device == self.device and t.is_contiguous()
                return t.to(self.device, non_blocking=True).contiguous(), not in_place
        first = arrays[0]
        if len(arrays) == 1 and isinstance(first, torch.Tensor) and first.is_cuda:
            t = first if first.dtype == dtype or (dtype == torch.int32 and first.dtype == torch.uint32) else first.to(dtype)
            if t.dtype == torch.uint32:
                t = t.view(torch.int32)
            in_place = t.data_ptr() == first.data_ptr() and t.is_contiguous()
            return t.contiguous().unsqueeze(0), not in_place
        shape = tuple(first.shape)
        out = torch.empty((len(arrays),) + shape, dtype=dtype, device=self.device)
        for i, a in enumerate(arrays):
            if isinstance(a, torch.Tensor):
                src = a
                if src.dtype == torch.uint32 and dtype == torch.int32:
                    src = src.view(torch.int32)
            else:
                a = np.asarray(a)
                if dtype == torch.int32 and a.dtype == np.uint32:
                    a = a.view(np.int32)
                elif dtype == torch.int32 and a.dtype != np.int32:
                    a = a.astype(np.uint32).view(np.int32)
                elif dtype == torch.float32 and a.dtype != np.float32:
                    a = a.astype(np.float32)
                src = torch.from_numpy(np.ascontiguousarray(a))
            out[i].copy_(src, non_blocking=True)
        return out, True

    def _jet_lut(self) -> torch.Tensor:
        """cv2.COLORMAP_JET as a [256,3] BGR table on the device (the table the reference applies, gcd.py:1702)."""
        lut = getattr(self, "_jet_lut_dev", None)
        if lut is None:
            import cv2

            table = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(-1, 1), cv2.COLORMAP_JET).reshape(256, 3)
            lut = torch.from_numpy(np.ascontiguousarray(table)).to(self.device)
            self._jet_lut_dev = lut
        return lut

    @staticmethod
    def _person_slots(t: FrameTables, num_people: int) -> List[int]:
        """Slot of person p = p-th ``human`` object in slot order (-1 when there are fewer)."""
        humans = [o.inst_idx for o in t.objects if o.class_name == "human"]
        return [humans[p] if p < len(humans) else -1 for p in range(num_people)]

    def _io_stream(self) -> torch.cuda.Stream:
        """One CUDA stream per I/O worker thread (device work issued from the file writers)."""
        import threading

        local = getattr(self, "_io_local", None)
        if local is None:
            local = self._io_local = threading.local()
        st = getattr(local, "stream", None)
        if st is None:
            st = local.stream = torch.cuda.Stream(device=self.device)
        return st

    # ------------------------------------------------------------------ device-formatted text files
    def _host_bytes(self, text: torch.Tensor, begin: int, end: int) -> memoryview:
        """text[begin:end] (device u8) through a grow-only pinned buffer; valid until the next call."""
        n = end - begin
        buf = getattr(self, "_text_host", None)
        if buf is None or buf.numel() < n:
            buf = torch.empty((max(n, 1 << 20),), dtype=torch.uint8, pin_memory=True)
            self._text_host = buf
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            buf[:n].copy_(text[begin:end], non_blocking=True)
            self.stream.synchronize()
        return memoryview(buf.numpy())[:n]

    def _write_depth_csv(self, labels: BatchLabels, chunk: int = 8) -> None:
        """gcd.py:1687-1688: depth/depth_%06d.csv, a few frames per formatting call."""
        depth = labels.device_outputs.get("depth")
        if depth is None:
            return
        ddir = os.path.join(self.output_dir, "depth")
        os.makedirs(ddir, exist_ok=True)
        B, H, W = depth.shape
        for f0 in range(0, B, chunk):
            nb = min(chunk, B - f0)
            with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
                chunk_depth = depth[f0:f0 + nb].reshape(nb * H, W)
                text, n_bytes, split = ops.format_fixed6(chunk_depth, split_rows=H)
                total = int(n_bytes.item())
                if total > text.numel():   # the 14-bytes-per-number estimate was short: format again into `total` bytes
                    text, n_bytes, split = ops.format_fixed6(chunk_depth, split_rows=H, capacity=total)
                offs = split.cpu().tolist() + [total]
            for k in range(nb):
                with open(os.path.join(ddir, f"depth_{labels.frame_ids[f0 + k]:06d}.csv"), "wb") as fh:
                    fh.write(self._host_bytes(text, offs[k], offs[k + 1]))

    def _write_pointclouds(self, labels: BatchLabels, chunk: int = 8) -> Optional[List[Optional[int]]]:
        """gcd.py:1716-1759: pointcloud/pointcloud_%06d.txt — from the frame's pointcloud annotator payload when it brought
        one (gcd.py:1720-1727), else from the depth map and the RGB image of the frame (gcd.py:1729-1759);
        returns the number of points per depth-derived cloud (None for the batch when there is nothing to write, None
        for a frame written from its annotator payload).  The clouds of ``chunk`` frames
        come out of one pair of launches (``cspe_depth_to_pointcloud_batch``), the text of each is formatted on the
        device and only its bytes cross PCIe."""
        depth = labels.device_outputs.get("depth")
        B = len(labels.frame_ids)
        # a frame that brought a pointcloud annotator payload is written from it (the reference's first choice,
        # gcd.py:1720-1727 -> save_pointcloud_with_rgb); only the others fall back to the depth map
        given = labels.pointcloud_rows if labels.pointcloud_rows is not None else [None] * B
        if depth is None and not any(g is not None for g in given):
            return None
        pdir = os.path.join(self.output_dir, "pointcloud")
        os.makedirs(pdir, exist_ok=True)
        rgbs = labels.rgb_images if labels.rgb_images is not None else [None] * B
        counts: List[int] = []
        for f0 in range(0, B, chunk):
            nb = min(chunk, B - f0)
            part = rgbs[f0:f0 + nb]
            with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
                d_rgb = None
                if depth is None or all(g is not None for g in given[f0:f0 + nb]):
                    clouds = [torch.empty((0, 6), dtype=torch.float64, device=self.device)] * nb   # nothing to derive
                else:
                    clouds = None
                if clouds is None and all(r is not None for r in part) and len({tuple(r.shape) for r in part}) == 1:
                    d_rgb = torch.stack([r.to(self.device) for r in part]).contiguous()
                if clouds is not None:
                    pass
                elif d_rgb is not None or all(r is None for r in part):
                    pts, offsets = ops.depth_to_pointcloud_batch(depth[f0:f0 + nb], d_rgb, labels.device_outputs["cam"][f0:f0 + nb])
                    offs = offsets.cpu().tolist()
                    clouds = [pts[offs[k]:offs[k + 1]] for k in range(nb)]
                else:   # frames with and without colour in one chunk: one frame at a time
                    clouds = []
                    for k in range(nb):
                        r = part[k].to(self.device).contiguous() if part[k] is not None else None
                        p1, n1 = ops.depth_to_pointcloud(depth[f0 + k], r, labels.device_outputs["cam"][f0 + k])
                        clouds.append(p1[: int(n1.item())])
                for k, cloud in enumerate(clouds):
                    if given[f0 + k] is not None:   # float64 copy: float32 coordinates print the same decimals
                        cloud = torch.from_numpy(np.ascontiguousarray(given[f0 + k], dtype=np.float64)).to(self.device)
                    points = int(cloud.shape[0])
                    # the reference's logger counts only clouds of the depth-map fallback (gcd.py:1754): None = not logged
                    counts.append(points if given[f0 + k] is None else None)
                    if points == 0:   # the reference saves nothing for an empty cloud (gcd.py:1749)
                        continue
                    text, n_bytes, _ = ops.format_fixed6(cloud, header="x y z r g b")
                    total = int(n_bytes.item())
                    if total > text.numel():
                        text, n_bytes, _ = ops.format_fixed6(cloud, header="x y z r g b", capacity=total)
                    with open(os.path.join(pdir, f"pointcloud_{labels.frame_ids[f0 + k]:06d}.txt"), "wb") as fh:
                        fh.write(self._host_bytes(text, 0, total))
        return counts

    def _write_frame_files(self, labels: BatchLabels, f: int, masks: Optional[List[ArrayLike]]) -> None:
        """The per-frame label files; runs on a worker thread (the native formatters, cv2.imwrite, np.save and
        file writes all release the GIL)."""
        fid = labels.frame_ids[f]
        ldir = os.path.join(self.output_dir, "labels")
        if "json" in self.formats:   # gcd.py:2071-2072
            with open(os.path.join(ldir, f"label_{fid:06d}.json"), "wb") as fh:
                fh.write(labels.label_json(f))
        if "yolo" in self.formats:
            with open(os.path.join(ldir, f"label_{fid:06d}.txt"), "wb") as fh:
                fh.write(labels.yolo_text(f))
        if "depth_png" in self.formats and labels.depth_image(f) is not None:
            import cv2

            cv2.imwrite(os.path.join(self.output_dir, "depth", f"depth_{fid:06d}.png"), labels.depth_image(f))   # gcd.py:1703-1704
        if "rgb_png" in self.formats and labels.rgb_images is not None and labels.rgb_images[f] is not None:
            import cv2

            rgb = labels.rgb_images[f]
            src = rgb if isinstance(rgb, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rgb))
            with torch.cuda.device(self.device), torch.cuda.stream(self._io_stream()):   # f4 kernel, gcd.py:1671
                bgr = ops.rgb_to_bgr(src.to(self.device, non_blocking=True).contiguous()).cpu().numpy()
            cv2.imwrite(os.path.join(self.output_dir, "rgb", f"rgb_{fid:06d}.png"), bgr)   # gcd.py:1672-1673
        if "mask" in self.formats and labels.device_outputs.get("mask") is not None:
            # the real mask where the reference saves a -1 placeholder (gcd.py:2066-2069), from the writer's own
            # device copy of the batch
            m = labels.device_outputs["mask"][f].cpu().numpy()
            np.save(os.path.join(ldir, f"instance_mask_{fid:06d}.npy"), m.astype(np.int32, copy=False))

    def _serialise(self, labels: BatchLabels, masks: Optional[List[ArrayLike]]) -> None:
        labels.synchronize()
        B = len(labels)
        if self.output_dir is not None:
            if "depth_png" in self.formats:
                os.makedirs(os.path.join(self.output_dir, "depth"), exist_ok=True)
            if "rgb_png" in self.formats:
                os.makedirs(os.path.join(self.output_dir, "rgb"), exist_ok=True)
            if "depth_csv" in self.formats:
                self._write_depth_csv(labels)
            if "yolo" in self.formats:
                labels.yolo_text(0)   # format the whole batch once, before the workers slice it
            if self.io_threads > 1 and B > 1:
                if self._io_pool is None:
                    from concurrent.futures import ThreadPoolExecutor

                    self._io_pool = ThreadPoolExecutor(max_workers=self.io_threads, thread_name_prefix="cspe-io")
                for fut in [self._io_pool.submit(self._write_frame_files, labels, f, masks) for f in range(B)]:
                    fut.result()   # re-raises a worker's exception here
            else:
                for f in range(B):
                    self._write_frame_files(labels, f, masks)
        if "coco" in self.formats:   # all annotations of the batch in one native call
            text, count = labels.coco_text(self._coco_annotation_count + 1)
            if count:
                self._coco_annotation_text.append(text)
                self._coco_annotation_count += count
        # bookkeeping stays on the caller's thread, in frame order
        cloud_points = None
        if "pointcloud" in self.formats and self.output_dir is not None:
            cloud_points = self._write_pointclouds(labels)
        for f in range(B):
            fid = labels.frame_ids[f]
            recs = labels.records(f)
            points = cloud_points[f] if cloud_points is not None else None
            if "coco" in self.formats:
                self._coco_images.append(formats.coco_image(fid, labels.width, labels.height, f"rgb_{fid:06d}.png"))
            dq = labels.depth_quality(f)
            if dq is not None:
                self.depth_quality_log.append({"frame_id": fid, "depth": dq})
            if self.quality is not None:   # the events of gcd.py:1567, 1684 / 1711, 2075, 2078
                self.quality.frame_start(fid, list(labels.camera_poses[f][:3]))
                self.quality.depth(dq, reason="annotator返回None或空")
                if points is not None:   # gcd.py:1754
                    self.quality.pointcloud(points > 0, points, "no valid depth pixel")
                self.quality.labels(len(recs))
                self.quality.frame_end(True)
            self.objects_total += len(recs)
            self.frames_written += 1
