"""Multi-GPU plumbing: one process per GPU (torchrun), model replicated, sentences / trials sharded by rank, and ONE
result gather -- the path has no reduction and no exchange step.  torch.distributed is used with the
"nccl" backend on GPUs (NVLink 5 / NVSwitch) and "gloo" in the CPU tests."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard(items: Sequence, rank: int, world: int) -> List:
    """Rank r of N takes items r::N (independent sentences / trials)."""
    return list(items[rank::world])


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition [lo, hi) of a sentence batch (keeps the gathered tensor in sentence order)."""
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def gather_token_ids(ys: torch.Tensor, world: int) -> torch.Tensor:
    """all_gather of the decoded ids [B/N, max_len] int64 -> [B, max_len] on every rank (<= 37 KB per rank at B = 64)."""
    if world == 1:
        return ys
    out = torch.empty((world * ys.shape[0],) + tuple(ys.shape[1:]), dtype=ys.dtype, device=ys.device)
    dist.all_gather_into_tensor(out, ys.contiguous())
    return out


def gather_records(records: List[tuple], world: int) -> List[tuple]:
    """Gather small per-trial records (trial_id, outcome, ...) from all ranks, sorted by trial id."""
    if world == 1:
        return sorted(records)
    parts = [None] * world
    dist.all_gather_object(parts, records)
    return sorted(x for part in parts for x in part)
