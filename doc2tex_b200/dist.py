"""Multi-GPU plumbing: independent batch shards per rank + ONE all-gather of the results.

The recognizer path shards naturally: every image is encoded and decoded independently
(SURVEY.md §8e — eval-mode BN is per sample, beams interact only within one image,
tools/beam.py:69-77), so ranks hold replicated weights, run their contiguous slice of the batch
with their own KV cache and CUDA graphs, and nothing crosses NVLink inside the decode loop.  The
only collective is a final all-gather of token ids / lengths / scores (~155 KB at B=256),
issued once per batch through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``n_items`` for ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(ids: torch.Tensor, lens: Optional[torch.Tensor] = None, scores: Optional[torch.Tensor] = None,
                   n_total: Optional[int] = None, group=None):
    """All-gather per-rank results into batch order.

    ids (b_r, T) int64, lens (b_r,) int32, scores (b_r,) fp32 — b_r may differ by one between ranks, so
    shards are padded to the largest and packed into ONE int64 buffer => one collective per batch.
    Returns (ids (n_total, T), lens or None, scores or None) on every rank.
    """
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ids, lens, scores
    world = dist.get_world_size(group)
    T = ids.shape[1]
    if n_total is None:
        cnt = torch.tensor([ids.shape[0]], device=ids.device, dtype=torch.int64)
        dist.all_reduce(cnt, group=group)
        n_total = int(cnt.item())
    b_max = (n_total + world - 1) // world
    packed = torch.zeros(b_max, T + 2, device=ids.device, dtype=torch.int64)
    b = ids.shape[0]
    packed[:b, :T] = ids
    if lens is not None:
        packed[:b, T] = lens.to(torch.int64)
    if scores is not None:
        packed[:b, T + 1] = scores.to(torch.float32).view(torch.int32).to(torch.int64)
    out = torch.empty(world * b_max, T + 2, device=ids.device, dtype=torch.int64)
    dist.all_gather_into_tensor(out, packed, group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        rows.append(out[r * b_max: r * b_max + (hi - lo)])
    full = torch.cat(rows, 0)
    g_ids = full[:, :T].contiguous()
    g_lens = full[:, T].to(torch.int32) if lens is not None else None
    g_scores = full[:, T + 1].to(torch.int32).view(torch.float32) if scores is not None else None
    return g_ids, g_lens, g_scores
