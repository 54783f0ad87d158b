"""Multi-GPU plumbing of the benchmark and tests: records are independent, so the path shards by
contiguous record ranges with no data-path collective. One process per GPU;
torch.distributed (nccl on GPUs, gloo in the CPU tests) is used only for barriers, the max-over-ranks
time and gathering the per-rank results, which are merged on the host in record order."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, 64-aligned (one flag word) range of records owned by `rank`."""
    per = (n_items + world - 1) // world
    per = (per + 63) // 64 * 64
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def merge_flags(parts: Sequence[Tuple[int, np.ndarray]], n_items: int) -> np.ndarray:
    """Merge per-shard flag bitmaps [(first_record, u64 words)] into one bitmap, in record order."""
    out = np.zeros((n_items + 63) // 64, dtype=np.uint64)
    for first, words in sorted(parts, key=lambda p: p[0]):
        assert first % 64 == 0
        w0 = first // 64
        out[w0:w0 + len(words)] |= words
    return out


def merge_hits(parts: Sequence[Tuple[int, np.ndarray]]) -> np.ndarray:
    """Concatenate per-shard sorted hit lists, rebasing the record index: the result is sorted."""
    outs = []
    for first, hits in sorted(parts, key=lambda p: p[0]):
        h = hits.copy()
        h["record"] += np.uint32(first)
        outs.append(h)
    return np.concatenate(outs) if outs else np.zeros(0)


class Dist:
    """Thin wrapper so that bench.py and the gloo tests share the reduction code."""

    def __init__(self, backend: str = "nccl", device=None):
        import torch.distributed as dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = dist
        self.device = device
        if self.world > 1 and not dist.is_initialized():
            kw = {}
            if backend == "nccl" and device is not None:
                kw["device_id"] = device
            dist.init_process_group(backend, **kw)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def _reduce(self, x: float, op) -> float:
        if self.world == 1:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def gather_objects(self, obj) -> List:
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.world > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()
