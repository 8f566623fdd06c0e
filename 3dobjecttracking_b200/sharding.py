"""Multi-GPU sharding of independent bodies.

Every Optimizer (rigid body) is independent within a frame (tracker.cpp:483-488), so bodies are partitioned
into contiguous blocks, one block per rank, models replicated, each rank receives only its bodies' frames.
There is NO collective on the iteration path; solved poses (12 floats per body) are all-gathered once per
frame for publishing / metrics.
"""
from __future__ import annotations


def shard_bounds(n_total: int, rank: int, world: int):
    """Contiguous block [first, first+count) of rank `rank`; blocks differ by at most one body."""
    if world <= 0 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def all_gather_poses(local_poses, counts=None, group=None):
    """All-gather [n_local,3,4] float32 pose tensors (CPU tensors with gloo, CUDA tensors with NCCL) into
    [n_total,3,4] in rank order. `counts`: per-rank body counts when the shards are uneven."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_poses
    world = dist.get_world_size(group)
    flat = local_poses.reshape(-1, 12).contiguous()
    if counts is None or len(set(counts)) == 1:
        out = torch.empty((world * flat.shape[0], 12), dtype=flat.dtype, device=flat.device)
        dist.all_gather_into_tensor(out, flat, group=group)
        return out.reshape(-1, 3, 4)
    n_max = max(counts)
    padded = torch.zeros((n_max, 12), dtype=flat.dtype, device=flat.device)
    padded[: flat.shape[0]] = flat
    out = torch.empty((world * n_max, 12), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * n_max: r * n_max + counts[r]] for r in range(world)]
    return torch.cat(parts, 0).reshape(-1, 3, 4)
