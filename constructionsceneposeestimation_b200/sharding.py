"""Frame-range sharding across ranks and the one collective of the path.

Frames (and the cameras of a rig, flattened to (frame, cam)) are independent, so each rank owns
a contiguous range and the data path has NO collective.  The only exchange is an all-gather
of the int64[10] per-class label histogram at the end of a sweep (80 bytes per rank): a
latency event on NVLink/NVSwitch, issued through torch.distributed (NCCL on GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ._lib import NUM_CLASSES


def frame_range(rank: int, world_size: int, num_frames: int) -> Tuple[int, int]:
    """[floor(r*F/W), floor((r+1)*F/W)) — contiguous, covers every frame exactly once."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    if num_frames < 0:
        raise ValueError("num_frames must be >= 0")
    return (rank * num_frames) // world_size, ((rank + 1) * num_frames) // world_size


def rig_frame_range(rank: int, world_size: int, num_rig_frames: int, cameras: int) -> List[Tuple[int, int]]:
    """(rig_frame, camera) pairs owned by ``rank`` when a multi-camera rig is flattened (config C4)."""
    lo, hi = frame_range(rank, world_size, num_rig_frames * cameras)
    return [(i // cameras, i % cameras) for i in range(lo, hi)]


def batches(lo: int, hi: int, batch: int) -> List[Tuple[int, int]]:
    """Split [lo, hi) into consecutive batches of at most ``batch`` frames."""
    if batch <= 0:
        raise ValueError("batch must be positive")
    return [(s, min(s + batch, hi)) for s in range(lo, hi, batch)]


def all_gather_histogram(hist: torch.Tensor) -> np.ndarray:
    """All-gather a per-rank int64[NUM_CLASSES] histogram -> int64 [world, NUM_CLASSES] on every rank.

    Works with whatever backend the default process group uses: NCCL needs the tensor on the
    rank's GPU (it already is), gloo needs CPU memory.
    """
    if tuple(hist.shape) != (NUM_CLASSES,) or hist.dtype != torch.int64:
        raise ValueError(f"histogram must be int64[{NUM_CLASSES}], got {hist.dtype}{tuple(hist.shape)}")
    world = dist.get_world_size()
    backend = dist.get_backend()
    src = hist if (backend == "nccl") == hist.is_cuda else (hist.cuda() if backend == "nccl" else hist.cpu())
    if backend == "nccl":
        out = torch.empty((world, NUM_CLASSES), dtype=torch.int64, device=src.device)
        dist.all_gather_into_tensor(out, src.contiguous())
    else:
        parts = [torch.empty_like(src) for _ in range(world)]
        dist.all_gather(parts, src.contiguous())
        out = torch.stack(parts)
    return out.cpu().numpy()
