"""Data-parallel plumbing for the multi-GPU path: samples are independent, so a global batch is split
contiguously across ranks (the in-process analogue of the reference's shell-level `--num-chunks/--chunk-idx`
sharding, scripts/v1_5/eval/cost_depth.sh:10-33 + eval/model_depth_loader.py:25-33) and the generated token ids are
joined with ONE all-gather.  No tensor other than the int32 ids ever crosses ranks.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.distributed as dist


def split_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) chunk of `n` samples for `rank` — same rule as the reference's get_chunk
    (model_depth_loader.py:25-33: chunk_size = ceil(n / k))."""
    chunk = math.ceil(n / world)
    b = min(n, rank * chunk)
    return b, min(n, b + chunk)


def shard_batch(world: int, rank: int, *tensors):
    """Slice dim 0 of every tensor (None passes through) to this rank's chunk."""
    n = next(t.shape[0] for t in tensors if t is not None)
    b, e = split_range(n, world, rank)
    return tuple(None if t is None else t[b:e] for t in tensors)


def all_gather_ids(local_ids: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """local_ids [B_local, N] int32/int64 on this rank's device -> [n_total, N] on every rank.
    One collective (NCCL over NVLink on GPUs, gloo in the CPU tests); chunks are padded to the common chunk size so a
    single all_gather_into_tensor suffices even when n_total % world != 0."""
    world = dist.get_world_size(group)
    chunk = math.ceil(n_total / world)
    N = local_ids.shape[1]
    pad = torch.zeros(chunk, N, dtype=local_ids.dtype, device=local_ids.device)
    pad[: local_ids.shape[0]] = local_ids
    out = torch.empty(world * chunk, N, dtype=local_ids.dtype, device=local_ids.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    rows: List[torch.Tensor] = []
    for r in range(world):
        b, e = split_range(n_total, world, r)
        rows.append(out[r * chunk: r * chunk + (e - b)])
    return torch.cat(rows, dim=0)
