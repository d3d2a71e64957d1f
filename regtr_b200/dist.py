"""Pair-level data parallelism: every pair is independent end to end, so ranks
take contiguous slices of the batch and the only collective is one all_gather of the poses
(6*B_local*12 floats).  Works with the nccl (GPU) and gloo (CPU tests) backends."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_pairs: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of `n_pairs` for `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_poses(pose_local: torch.Tensor, n_pairs: int):
    """pose_local (L, B_local, 3, 4) -> (L, n_pairs, 3, 4) on every rank, in global pair order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return pose_local
    world, rank = dist.get_world_size(), dist.get_rank()
    L = pose_local.shape[0]
    b_max = -(-n_pairs // world)
    buf = pose_local.new_zeros((L, b_max, 3, 4))
    buf[:, :pose_local.shape[1]] = pose_local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_pairs, r, world)
        parts.append(out[r][:, :hi - lo])
    return torch.cat(parts, dim=1)
