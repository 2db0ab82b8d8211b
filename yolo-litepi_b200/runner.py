"""Frame-sharded multi-GPU execution.

The path shards on frames: no cross-frame state, weights replicated, so rank r of W
(one process per GPU) takes frames ``i % W == r`` and there is NO data-path collective.
The single exchange is the final gather of packed detection records
(9 x int32 each: frame_id, box f32 x4, det_conf, det_cls, cls_cls, cls_conf) --
a counts all-gather followed by a padded payload all-gather (NCCL on GPUs, gloo in the
CPU tests).  The reference is single-process and has nothing comparable.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist

REC_WORDS = 9


def shard_indices(n_frames: int, rank: int, world: int) -> List[int]:
    """Global frame ids owned by ``rank``: i with i % world == rank."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    return list(range(rank, n_frames, world))


def gather_records(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather variable-length record tensors [n_r, 9] int32 -> [sum n_r, 9], rank order.
    Works on whatever device ``local`` lives on (cuda -> NCCL, cpu -> gloo)."""
    if local.dim() != 2 or local.shape[1] != REC_WORDS or local.dtype != torch.int32:
        raise ValueError("records must be [n, 9] int32")
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    sizes = [int(c.item()) for c in cnts]
    mx = max(sizes)
    if mx == 0:
        return local
    padded = torch.zeros((mx, REC_WORDS), dtype=torch.int32, device=local.device)
    padded[:local.shape[0]] = local
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def sort_records(rec: np.ndarray) -> np.ndarray:
    """Stable order by frame id (ranks interleave frames), keeping per-frame detection order."""
    if rec.shape[0] == 0:
        return rec
    return rec[np.argsort(rec[:, 0], kind="stable")]


def run_sharded(pipeline, frames: Sequence[np.ndarray], conf: float, iou: float, min_area: int,
                rank: int, world: int, group=None) -> np.ndarray:
    """Process this rank's share of ``frames`` (global list) in batches and gather every rank's
    records; returns the global record array sorted by frame id (same on every rank)."""
    from .detector import FrameBatch
    mine = shard_indices(len(frames), rank, world)
    chunks = []
    B = pipeline.max_batch
    for i in range(0, len(mine), B):
        ids = mine[i:i + B]
        fb = FrameBatch.from_host([frames[j] for j in ids], pipeline.device)
        fid = torch.tensor(ids, dtype=torch.int32, device=pipeline.device)
        n = pipeline.run_device(fb, conf, iou, min_area, fid)
        chunks.append(pipeline.records[:n].clone())
    local = torch.cat(chunks) if chunks else torch.zeros((0, REC_WORDS), dtype=torch.int32, device=pipeline.device)
    allrec = gather_records(local, group)
    return sort_records(allrec.cpu().numpy())
