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


def init_process_group(backend: str = None) -> None:
    """One process per GPU (torchrun): NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests.  No-op for a single process."""
    rank, world, local_rank = rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)


def destroy_process_group() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values: Sequence[float], device=None) -> List[float]:
    """Element-wise MAX of a few host scalars over the ranks (timings: a multi-GPU number is the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def sum_over_ranks(values: Sequence[float], device=None) -> List[float]:
    if not (dist.is_available() and dist.is_initialized()):
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


class OverlappedGather:
    """The path's only collective -- the gather of the decoded token ids -- taken OFF the critical path: the all_gather of batch i runs
    on a side stream while the GPU already decodes batch i+1 (it is 36 KB per rank: pure launch + NCCL latency, ~0.25 ms, which cost
    2 % of scaling efficiency at 8 GPUs when it sat serially after every decode).  `submit(ys)` is asynchronous with respect to the
    compute stream; `result()` waits for the last submitted gather and returns the [world * B, max_len] tensor."""

    def __init__(self, world: int, device):
        self.world = world
        self.stream = torch.cuda.Stream(device=device) if (world > 1 and torch.cuda.is_available()) else None
        self.out = None
        self.src = None
        self.done = None

    def submit(self, ys: torch.Tensor) -> None:
        if self.world == 1:
            self.out = ys
            return
        if self.out is None or self.out.shape[1:] != ys.shape[1:] or self.out.shape[0] != self.world * ys.shape[0]:
            self.out = torch.empty((self.world * ys.shape[0],) + tuple(ys.shape[1:]), dtype=ys.dtype, device=ys.device)
        ready = torch.cuda.Event()
        ready.record()                                   # ys is complete at this point of the compute stream
        self.src = ys                                    # keep the source alive until the side stream has read it
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            dist.all_gather_into_tensor(self.out, ys.contiguous())
            self.done = torch.cuda.Event()
            self.done.record()

    def result(self) -> torch.Tensor:
        if self.world > 1 and self.done is not None:
            torch.cuda.current_stream().wait_event(self.done)
        return self.out
