"""Multi-GPU plumbing for the sharded single-dispatch mode: torch.distributed only.

One huge dispatch: every rank holds both meshes and both BVHs and traverses its own slice of the query leaves
(mcb200_result_set_shard); the pair and registry buffers are then exchanged with an all-gather.  NCCL has no
all-gather-v, so counts travel first and payloads are padded to the largest count.  The merged pair set is sorted, which
makes the result independent of the number of ranks.  Works on any backend (gloo on CPU for the tests, nccl on GPUs).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_of_leaf(first_leaf: int, nparts: int, chunk: int = 4096) -> int:
    """Which rank traverses the query group that starts at sorted leaf `first_leaf` (traverse.cu: k_traverse)."""
    return (first_leaf // chunk) % nparts


def allgatherv(t: torch.Tensor) -> Tuple[torch.Tensor, List[int]]:
    """All-gather of 1-D tensors of different lengths.  Returns (concatenation in rank order, per-rank counts)."""
    world = dist.get_world_size()
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    counts = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(counts, n)
    counts_h = [int(c.item()) for c in counts]
    m = max(max(counts_h), 1)
    padded = torch.zeros(m, dtype=t.dtype, device=t.device)
    padded[:t.numel()] = t
    outs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(outs, padded)
    return torch.cat([o[:c] for o, c in zip(outs, counts_h)]), counts_h


def merge_pairs(local_pairs: torch.Tensor) -> torch.Tensor:
    """Union of the ranks' candidate pairs (int64 view of src << 32 | cut), ascending — the single-GPU output."""
    allp, _ = allgatherv(local_pairs)
    return torch.sort(allp).values


def merge_records(local_records: torch.Tensor) -> torch.Tensor:
    """Union of the ranks' registry records ([n, 4] float64 rows whose first 8 bytes are (edge, face)), ordered by
    (edge, face) like the single-GPU output."""
    flat, counts = allgatherv(local_records.reshape(-1))
    rec = flat.reshape(-1, 4)
    if rec.shape[0] == 0:
        return rec
    ids = rec[:, 0].contiguous().view(torch.int32).reshape(-1, 2).to(torch.int64) & 0xFFFFFFFF  # (edge, face)
    key = (ids[:, 0] << 32) | ids[:, 1]
    return rec[torch.argsort(key)]
