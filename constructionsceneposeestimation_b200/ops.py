"""Device-tensor wrappers over the libcspe C ABI.

torch is plumbing here (device memory + streams); every function forwards raw device pointers
and the current CUDA stream to the hand-written kernels and never falls back to torch math.
All calls are asynchronous with respect to the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import CAM_STRIDE, NUM_CLASSES, POSE_STRIDE, SCAN_FIELDS


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def as_u32_mask(mask: torch.Tensor) -> torch.Tensor:
    """Validate an id mask: uint32 (Replicator native) or int32 with the same bits; no copy."""
    if mask.dtype not in (torch.uint32, torch.int32):
        raise ValueError(f"instance mask must be uint32 or int32, got {mask.dtype}")
    if not mask.is_cuda:
        raise ValueError("mask must be a CUDA tensor (no CPU fallback exists)")
    if not mask.is_contiguous():
        raise ValueError("mask must be contiguous")
    return mask


def mask_scan(mask: torch.Tensor, id2slot: torch.Tensor, num_slots: int, out: Optional[torch.Tensor] = None,
              accumulate: bool = False) -> torch.Tensor:
    """K1: mask uint32/int32 [B,H,W] (or [H,W]) + LUT int32 [L] or [B,L] -> int32 [B,N,5].
    ``accumulate=True`` merges into an ``out`` that already holds scan entries (no init launch)."""
    lib = _lib.load()
    mask = as_u32_mask(mask)
    if mask.dim() == 2:
        mask = mask.unsqueeze(0)
    _dev(id2slot, torch.int32, "id2slot")
    B, H, W = mask.shape
    if id2slot.dim() == 1:
        lut_len, lut_stride = id2slot.shape[0], 0
    else:
        if id2slot.shape[0] != B:
            raise ValueError(f"id2slot has {id2slot.shape[0]} rows for a batch of {B}")
        lut_len, lut_stride = id2slot.shape[1], id2slot.shape[1]
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs an initialised `out`")
        out = torch.empty((B, num_slots, SCAN_FIELDS), dtype=torch.int32, device=mask.device)
    else:
        _dev(out, torch.int32, "out")
        if tuple(out.shape) != (B, num_slots, SCAN_FIELDS):
            raise ValueError(f"out has shape {tuple(out.shape)}, expected {(B, num_slots, SCAN_FIELDS)}")
    fn = lib.cspe_mask_scan_accumulate if accumulate else lib.cspe_mask_scan
    with torch.cuda.device(mask.device):
        rc = fn(mask.data_ptr(), B, H, W, id2slot.data_ptr(), lut_len, lut_stride, num_slots,
                                out.data_ptr(), _stream_ptr())
    _lib.check("cspe_mask_scan", rc)
    return out


def mask_scan_depth_stats(mask: torch.Tensor, depth: torch.Tensor, id2slot: torch.Tensor, num_slots: int
                          ) -> Tuple[torch.Tensor, torch.Tensor]:
    """K1 + fused depth statistics.  Returns (scan int32 [B,N,5], stats uint8 [B,48])."""
    lib = _lib.load()
    mask = as_u32_mask(mask)
    if mask.dim() == 2:
        mask, depth = mask.unsqueeze(0), depth.unsqueeze(0)
    _dev(depth, torch.float32, "depth")
    _dev(id2slot, torch.int32, "id2slot")
    if depth.shape != mask.shape:
        raise ValueError(f"depth {tuple(depth.shape)} and mask {tuple(mask.shape)} differ")
    B, H, W = mask.shape
    lut_len = id2slot.shape[-1]
    lut_stride = 0 if id2slot.dim() == 1 else lut_len
    out = torch.empty((B, num_slots, SCAN_FIELDS), dtype=torch.int32, device=mask.device)
    stats = torch.empty((B, _lib.DEPTH_STATS_DTYPE.itemsize), dtype=torch.uint8, device=mask.device)
    with torch.cuda.device(mask.device):
        rc = lib.cspe_mask_scan_depth_stats(mask.data_ptr(), depth.data_ptr(), B, H, W, id2slot.data_ptr(), lut_len,
                                            lut_stride, num_slots, out.data_ptr(), stats.data_ptr(), _stream_ptr())
    _lib.check("cspe_mask_scan_depth_stats", rc)
    return out, stats


def depth_stats(depth: torch.Tensor) -> torch.Tensor:
    """f2: depth float32 [B,H,W] -> stats uint8 [B,48] (view with DEPTH_STATS_DTYPE on host)."""
    lib = _lib.load()
    if depth.dim() == 2:
        depth = depth.unsqueeze(0)
    _dev(depth, torch.float32, "depth")
    B, H, W = depth.shape
    stats = torch.empty((B, _lib.DEPTH_STATS_DTYPE.itemsize), dtype=torch.uint8, device=depth.device)
    with torch.cuda.device(depth.device):
        rc = lib.cspe_depth_stats(depth.data_ptr(), B, H, W, stats.data_ptr(), _stream_ptr())
    _lib.check("cspe_depth_stats", rc)
    return stats


def union_records(records: torch.Tensor, base: int, offsets: torch.Tensor, members: torch.Tensor) -> torch.Tensor:
    """Object-level records for multi-mesh objects (``record_fallback="union"``): records uint8 [B,R,stride] with
    ``R >= base + U``; offsets int32 [U+1] or [B,U+1], members int32 [M] or [B,M] (record indices < base).  Writes
    records[:, base + u] in place (world-axis-aligned range of the members' corners, identity transform) and returns
    ``records``."""
    lib = _lib.load()
    _dev(records, torch.uint8, "records")
    _dev(offsets, torch.int32, "offsets")
    _dev(members, torch.int32, "members")
    if records.dim() != 3:
        raise ValueError(f"records must be uint8 [B,R,stride], got {tuple(records.shape)}")
    B, R, stride = records.shape
    if offsets.dim() != members.dim() or offsets.dim() not in (1, 2) or (offsets.dim() == 2 and
                                                                         (offsets.shape[0] != B or members.shape[0] != B)):
        raise ValueError(f"offsets {tuple(offsets.shape)} / members {tuple(members.shape)}: need [U+1] / [M] or [B,U+1] / [B,M]")
    U = offsets.shape[-1] - 1
    if U < 0 or base < 0 or base + U > R:
        raise ValueError(f"base {base} + {U} union objects exceed the {R} records per frame")
    off_stride = offsets.shape[1] if offsets.dim() == 2 else 0
    mem_stride = members.shape[1] if members.dim() == 2 else 0
    with torch.cuda.device(records.device):
        rc = lib.cspe_union_records(records.data_ptr(), stride, R, base, offsets.data_ptr(), off_stride,
                                    members.data_ptr(), mem_stride, B, U, _stream_ptr())
    _lib.check("cspe_union_records", rc)
    return records


def project_objects(records: torch.Tensor, obj_record: torch.Tensor, cam: torch.Tensor):
    """K2: records uint8 [B,R,stride], obj_record int32 [B,N], cam f64 [B,24]
    -> uv f64 [B,N,8,2], z f64 [B,N,8], pose f64 [B,N,16], loose f64 [B,N,4], flags u8 [B,N]."""
    lib = _lib.load()
    _dev(records, torch.uint8, "records")
    _dev(obj_record, torch.int32, "obj_record")
    _dev(cam, torch.float64, "cam")
    B, N = obj_record.shape
    if records.dim() != 3 or records.shape[0] != B:
        raise ValueError(f"records must be uint8 [B,R,stride] with B={B}, got {tuple(records.shape)}")
    if tuple(cam.shape) != (B, CAM_STRIDE):
        raise ValueError(f"cam must be [{B},{CAM_STRIDE}], got {tuple(cam.shape)}")
    R, stride = records.shape[1], records.shape[2]
    dev = obj_record.device
    uv = torch.empty((B, N, 8, 2), dtype=torch.float64, device=dev)
    z = torch.empty((B, N, 8), dtype=torch.float64, device=dev)
    pose = torch.empty((B, N, POSE_STRIDE), dtype=torch.float64, device=dev)
    loose = torch.empty((B, N, 4), dtype=torch.float64, device=dev)
    flags = torch.empty((B, N), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cspe_project_objects(records.data_ptr(), stride, R, obj_record.data_ptr(), cam.data_ptr(), B, N,
                                      uv.data_ptr(), z.data_ptr(), pose.data_ptr(), loose.data_ptr(),
                                      flags.data_ptr(), _stream_ptr())
    _lib.check("cspe_project_objects", rc)
    return uv, z, pose, loose, flags


def keypoints(joints: torch.Tensor, depth: torch.Tensor, cam: torch.Tensor, tol: float = 0.15):
    """K3: joints f32 [B,P,J,3], depth f32 [B,H,W], cam f64 [B,24]
    -> kp f64 [B,P,J,2], kz f64 [B,P,J], vis u8 [B,P,J]."""
    lib = _lib.load()
    _dev(joints, torch.float32, "joints")
    _dev(depth, torch.float32, "depth")
    _dev(cam, torch.float64, "cam")
    B, P, J, three = joints.shape
    if three != 3:
        raise ValueError("joints must be [B,P,J,3]")
    if depth.dim() != 3 or depth.shape[0] != B:
        raise ValueError(f"depth must be [B,H,W] with B={B}, got {tuple(depth.shape)}")
    if tuple(cam.shape) != (B, CAM_STRIDE):
        raise ValueError(f"cam must be [{B},{CAM_STRIDE}], got {tuple(cam.shape)}")
    H, W = depth.shape[1], depth.shape[2]
    dev = joints.device
    kp = torch.empty((B, P, J, 2), dtype=torch.float64, device=dev)
    kz = torch.empty((B, P, J), dtype=torch.float64, device=dev)
    vis = torch.empty((B, P, J), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cspe_keypoints(joints.data_ptr(), B, P, J, depth.data_ptr(), H, W, cam.data_ptr(), float(tol),
                                kp.data_ptr(), kz.data_ptr(), vis.data_ptr(), _stream_ptr())
    _lib.check("cspe_keypoints", rc)
    return kp, kz, vis


def emit(scan, uv, z, pose, loose, flags, slot_class, H: int, W: int, min_pixels: int = 1, frame_base: int = 0,
         class_hist: Optional[torch.Tensor] = None, records: Optional[torch.Tensor] = None,
         n_out: Optional[torch.Tensor] = None):
    """K4: -> records uint8 [B,N,408] (view with RECORD_DTYPE on host), n_out int32 [B],
    class_hist int64 [10] (accumulated into when passed in)."""
    lib = _lib.load()
    _dev(scan, torch.int32, "scan")
    _dev(slot_class, torch.int32, "slot_class")
    _dev(flags, torch.uint8, "flags")
    for name, t in (("uv", uv), ("z", z), ("pose", pose), ("loose", loose)):
        _dev(t, torch.float64, name)
    B, N = slot_class.shape
    dev = scan.device
    if class_hist is None:
        class_hist = torch.zeros((NUM_CLASSES,), dtype=torch.int64, device=dev)
    else:
        _dev(class_hist, torch.int64, "class_hist")
    if records is None:
        records = torch.empty((B, N, _lib.RECORD_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    if n_out is None:
        n_out = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cspe_emit(scan.data_ptr(), uv.data_ptr(), z.data_ptr(), pose.data_ptr(), loose.data_ptr(),
                           flags.data_ptr(), slot_class.data_ptr(), B, N, H, W, min_pixels, frame_base,
                           records.data_ptr(), n_out.data_ptr(), class_hist.data_ptr(), _stream_ptr())
    _lib.check("cspe_emit", rc)
    return records, n_out, class_hist


def format_yolo(records: torch.Tensor, n_out: torch.Tensor, frame_stride: Optional[int] = None):
    """S6 / f3 on the device: records u8 [B,N,408] + n_out int32 [B] (K4's outputs) -> (text u8 [B, frame_stride],
    n_bytes int32 [B]); frame f's YOLO label file is ``text[f, :n_bytes[f]]`` (``class cx cy w h`` lines, six
    decimals).  n_bytes[f] > frame_stride means the frame's text was cut, -1 an unprintable box value."""
    lib = _lib.load()
    _dev(records, torch.uint8, "records")
    _dev(n_out, torch.int32, "n_out")
    if records.dim() != 3 or records.shape[2] != _lib.RECORD_DTYPE.itemsize or records.shape[0] != n_out.shape[0]:
        raise ValueError(f"records must be u8 [B,N,{_lib.RECORD_DTYPE.itemsize}] with B = len(n_out), got {tuple(records.shape)}")
    B, N = records.shape[0], records.shape[1]
    stride = 48 * N if frame_stride is None else int(frame_stride)
    text = torch.empty((B, stride), dtype=torch.uint8, device=records.device)
    n_bytes = torch.empty((B,), dtype=torch.int32, device=records.device)
    with torch.cuda.device(records.device):
        rc = lib.cspe_format_yolo(records.data_ptr(), n_out.data_ptr(), B, N, text.data_ptr(), stride,
                                  n_bytes.data_ptr(), _stream_ptr())
    _lib.check("cspe_format_yolo", rc)
    return text, n_bytes


def format_coco(records: torch.Tensor, n_out: torch.Tensor, ann_state: torch.Tensor, frame_stride: Optional[int] = None):
    """S6 / f3 on the device: records u8 [B,N,408] + n_out int32 [B] -> (text u8 [B, frame_stride], n_bytes int32 [B]);
    frame f's COCO annotation objects, ", "-joined, are ``text[f, :n_bytes[f]]`` (join the frames with ", " as well).
    ``ann_state`` int64 [2] on the device carries the running annotation count across calls (zero it at sweep start)."""
    lib = _lib.load()
    _dev(records, torch.uint8, "records")
    _dev(n_out, torch.int32, "n_out")
    _dev(ann_state, torch.int64, "ann_state")
    if records.dim() != 3 or records.shape[2] != _lib.RECORD_DTYPE.itemsize or records.shape[0] != n_out.shape[0]:
        raise ValueError(f"records must be u8 [B,N,{_lib.RECORD_DTYPE.itemsize}] with B = len(n_out), got {tuple(records.shape)}")
    if ann_state.numel() != 2:
        raise ValueError("ann_state must be int64 [2]")
    B, N = records.shape[0], records.shape[1]
    stride = 224 * N if frame_stride is None else int(frame_stride)
    text = torch.empty((B, stride), dtype=torch.uint8, device=records.device)
    n_bytes = torch.empty((B,), dtype=torch.int32, device=records.device)
    with torch.cuda.device(records.device):
        rc = lib.cspe_format_coco(records.data_ptr(), n_out.data_ptr(), B, N, ann_state.data_ptr(), text.data_ptr(), stride,
                                  n_bytes.data_ptr(), _stream_ptr())
    _lib.check("cspe_format_coco", rc)
    return text, n_bytes


def depth_to_pointcloud(depth: torch.Tensor, rgb: Optional[torch.Tensor], cam: torch.Tensor,
                        capacity: Optional[int] = None):
    """f1: depth f32 [H,W], rgb u8 [H,W,C>=3] or None, cam f64 [24] -> (points f64 [cap,6], n int64 [1])."""
    lib = _lib.load()
    _dev(depth, torch.float32, "depth")
    _dev(cam, torch.float64, "cam")
    H, W = depth.shape
    ch = 0
    if rgb is not None:
        _dev(rgb, torch.uint8, "rgb")
        if rgb.dim() != 3 or rgb.shape[0] != H or rgb.shape[1] != W:
            raise ValueError(f"rgb must be [H,W,C], got {tuple(rgb.shape)}")
        ch = rgb.shape[2]
    if capacity is None:
        capacity = H * W
    dev = depth.device
    out = torch.empty((capacity, 6), dtype=torch.float64, device=dev)
    n = torch.empty((1,), dtype=torch.int64, device=dev)
    ws_bytes = lib.cspe_pointcloud_workspace_bytes(H, W)
    ws = torch.empty(((ws_bytes + 7) // 8,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cspe_depth_to_pointcloud(depth.data_ptr(), _ptr(rgb), ch, H, W, cam.data_ptr(), out.data_ptr(),
                                          capacity, n.data_ptr(), ws.data_ptr(), _stream_ptr())
    _lib.check("cspe_depth_to_pointcloud", rc)
    return out, n


def depth_to_pointcloud_batch(depth: torch.Tensor, rgb: Optional[torch.Tensor], cam: torch.Tensor,
                              capacity: Optional[int] = None):
    """f1 for a batch: depth f32 [B,H,W], rgb u8 [B,H,W,C>=3] or None, cam f64 [B,24] -> (points f64 [cap,6],
    offsets int64 [B+1]); frame f's cloud is ``points[offsets[f]:offsets[f+1]]`` (row-major pixel order), two
    launches for the whole batch."""
    lib = _lib.load()
    _dev(depth, torch.float32, "depth")
    _dev(cam, torch.float64, "cam")
    if depth.dim() != 3:
        raise ValueError(f"depth must be [B,H,W], got {tuple(depth.shape)}")
    B, H, W = depth.shape
    if tuple(cam.shape) != (B, CAM_STRIDE):
        raise ValueError(f"cam must be [{B},{CAM_STRIDE}], got {tuple(cam.shape)}")
    ch = 0
    if rgb is not None:
        _dev(rgb, torch.uint8, "rgb")
        if rgb.dim() != 4 or tuple(rgb.shape[:3]) != (B, H, W):
            raise ValueError(f"rgb must be [B,H,W,C], got {tuple(rgb.shape)}")
        ch = rgb.shape[3]
    if capacity is None:
        capacity = B * H * W
    dev = depth.device
    out = torch.empty((capacity, 6), dtype=torch.float64, device=dev)
    offsets = torch.empty((B + 1,), dtype=torch.int64, device=dev)
    ws_bytes = lib.cspe_pointcloud_batch_workspace_bytes(B, H, W)
    ws = torch.empty(((ws_bytes + 7) // 8,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cspe_depth_to_pointcloud_batch(depth.data_ptr(), _ptr(rgb), ch, B, H, W, cam.data_ptr(), out.data_ptr(),
                                                capacity, offsets.data_ptr(), ws.data_ptr(), _stream_ptr())
    _lib.check("cspe_depth_to_pointcloud_batch", rc)
    return out, offsets


def format_fixed6(values: torch.Tensor, n_rows: Optional[torch.Tensor] = None, header: Optional[str] = None,
                  split_rows: int = 0, capacity: Optional[int] = None, out: Optional[torch.Tensor] = None):
    """f3: the bytes ``np.savetxt(f, values, fmt='%.6f', delimiter=' ', header=header, comments='')`` writes
    (gcd.py:1688, 1752), formatted on the device.

    values f32 / f64 [rows, cols]; n_rows int64 [1] on the device = live rows (None = all);
    split_rows > 0 additionally returns the byte offset of every ``split_rows``-th row.
    Returns (text u8 [capacity], n_bytes int64 [1], split_offsets int64 [ceil(rows/split_rows)] or None);
    ``n_bytes`` is the size of the complete text — more than ``capacity`` means it was cut (call again with
    that capacity), -1 means a finite value of magnitude >= 2^128 was met."""
    lib = _lib.load()
    if values.dim() != 2:
        raise ValueError(f"values must be [rows, cols], got {tuple(values.shape)}")
    if values.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"values must be float32 or float64, got {values.dtype}")
    _dev(values, values.dtype, "values")
    rows, cols = values.shape
    if n_rows is not None:
        _dev(n_rows, torch.int64, "n_rows")
    dev = values.device
    if out is not None:
        _dev(out, torch.uint8, "out")
        capacity = out.numel()
    elif capacity is None:
        # typical line: "-123.456789 " = 12 bytes per value; the caller retries with n_bytes if that was short
        capacity = rows * cols * 14 + 64
    if out is None:
        out = torch.empty((capacity,), dtype=torch.uint8, device=dev)
    n_bytes = torch.empty((1,), dtype=torch.int64, device=dev)
    split = None
    if split_rows > 0:
        split = torch.zeros(((rows + split_rows - 1) // split_rows,), dtype=torch.int64, device=dev)
    ws_bytes = lib.cspe_text_workspace_bytes(rows, cols)
    ws = torch.empty(((ws_bytes + 7) // 8,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cspe_format_fixed6(values.data_ptr(), 1 if values.dtype == torch.float64 else 0, rows, _ptr(n_rows),
                                    cols, header.encode("utf-8") if header is not None else None, out.data_ptr(),
                                    capacity, n_bytes.data_ptr(), split_rows, _ptr(split), ws.data_ptr(), _stream_ptr())
    _lib.check("cspe_format_fixed6", rc)
    return out, n_bytes, split


def savetxt_bytes(values: torch.Tensor, n_rows: Optional[torch.Tensor] = None, header: Optional[str] = None) -> bytes:
    """Complete ``np.savetxt(fmt='%.6f', delimiter=' ')`` text of a device matrix as host bytes (synchronises)."""
    text, n_bytes, _ = format_fixed6(values, n_rows, header)
    need = int(n_bytes.item())
    if need < 0:
        raise ValueError("format_fixed6: a finite value of magnitude >= 2^128 cannot be formatted")
    if need > text.numel():
        text, n_bytes, _ = format_fixed6(values, n_rows, header, capacity=need)
    return text[:need].cpu().numpy().tobytes()


def depth_colormap(depth: torch.Tensor, lut_bgr: torch.Tensor, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """f4: depth f32 [B,H,W] + colour LUT u8 [256,3] (BGR) -> u8 [B,H,W,3]; stats from depth_stats if not given."""
    lib = _lib.load()
    if depth.dim() == 2:
        depth = depth.unsqueeze(0)
    _dev(depth, torch.float32, "depth")
    _dev(lut_bgr, torch.uint8, "lut_bgr")
    if tuple(lut_bgr.shape) != (256, 3):
        raise ValueError(f"lut_bgr must be [256,3], got {tuple(lut_bgr.shape)}")
    if stats is None:
        stats = depth_stats(depth)
    B, H, W = depth.shape
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=depth.device)
    with torch.cuda.device(depth.device):
        rc = lib.cspe_depth_colormap(depth.data_ptr(), B, H, W, stats.data_ptr(), lut_bgr.data_ptr(), out.data_ptr(),
                                     _stream_ptr())
    _lib.check("cspe_depth_colormap", rc)
    return out


def rgb_to_bgr(rgb: torch.Tensor) -> torch.Tensor:
    """f4: u8 [..., C>=3] -> u8 [..., 3] with channels reversed (alpha dropped)."""
    lib = _lib.load()
    _dev(rgb, torch.uint8, "rgb")
    ch = rgb.shape[-1]
    n = rgb.numel() // max(ch, 1)
    out = torch.empty(tuple(rgb.shape[:-1]) + (3,), dtype=torch.uint8, device=rgb.device)
    with torch.cuda.device(rgb.device):
        rc = lib.cspe_rgb_to_bgr(rgb.data_ptr(), ch, n, out.data_ptr(), _stream_ptr())
    _lib.check("cspe_rgb_to_bgr", rc)
    return out


def device_info() -> Tuple[int, int, int]:
    lib = _lib.load()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _lib.check("cspe_device_info", lib.cspe_device_info(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value
