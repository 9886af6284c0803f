"""Multi-GPU = independent replicas over contiguous batch chunks; no collective inside the data path.

Samples never interact (per-sample attention, per-sample top-k, per-sample KV), so the path shards exactly like the
reference's launcher: one process per GPU, chunk = ceil(len / n) contiguous items
(llava/dynamic_eval/model_vqa_loader.py:28-38; run/dynamic_eval/eval_for_vqav2.sh:3-33, results merged by `cat`).
Here the merge is one `all_gather` of the generated ids AFTER the timed region (NCCL over NVLink on the B200 box,
gloo in the CPU tests) plus a MAX all-reduce of the elapsed time.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
import torch.distributed as dist


def chunk_bounds(n_items: int, n_chunks: int, k: int):
    """[lo, hi) of chunk k under the reference's `split_list` rule (chunk_size = ceil(n / n_chunks))."""
    size = math.ceil(n_items / n_chunks) if n_items else 0
    lo = min(k * size, n_items)
    return lo, min(lo + size, n_items)


def get_chunk(lst: Sequence, n: int, k: int):
    lo, hi = chunk_bounds(len(lst), n, k)
    return lst[lo:hi]


def gather_token_ids(local: torch.Tensor, n_items: int) -> torch.Tensor:
    """local int64 [n_local, n_new] -> [n_items, n_new] on every rank, in global sample order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    size = math.ceil(n_items / world)
    pad = torch.zeros(size, local.shape[1], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = []
    for r in range(world):
        lo, hi = chunk_bounds(n_items, world, r)
        out.append(bufs[r][: hi - lo])
    return torch.cat(out, 0)


def max_over_ranks(x: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
