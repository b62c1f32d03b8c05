"""Batch sharding and the path's single collective.

Images are independent units through encode -> VQ -> decode (no cross-sample op on the path), so
the work is sharded contiguously over ranks with replicated weights and NO data-path collective.
The only exchange is an all-reduce(sum) of the codebook-usage histogram (int64[n_e]) and of
(sum of squared errors, element count), so that the global commitment loss (1+beta)*SSE/count and
the global usage histogram are exact — identical to a single-GPU run over the whole set.  One
process per GPU; torch.distributed carries it (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `total` units for `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_usage(hist: torch.Tensor, sums: torch.Tensor, group=None):
    """In-place all-reduce(sum) of hist (int64 [n_e]) and sums (float64 [2] = SSE, count)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return hist, sums


def global_loss(sums: torch.Tensor, beta: float = 0.25) -> float:
    """(1 + beta) * SSE / count  (reference stage1/quantize.py:33 evaluated over the whole sharded set)."""
    return float((1.0 + beta) * sums[0] / sums[1])


def tokenize_sharded(model, images_fn, total: int, batch: int, group=None):
    """Tokenize `total` images sharded over the ranks of `group`.

    images_fn(lo, hi) -> CUDA tensor [hi-lo, 3, H, W] for global image indices [lo, hi) (content must
    depend only on the global index so results are independent of the world size).
    Returns (indices of this rank's shard [n, N] int64, global histogram, global loss).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(total, rank, world)
    n_e = model.quantize.n_e
    dev = next(model.parameters()).device
    hist = torch.zeros(n_e, device=dev, dtype=torch.int64)
    sums = torch.zeros(2, device=dev, dtype=torch.float64)
    out = []
    for b0 in range(lo, hi, batch):
        b1 = min(b0 + batch, hi)
        z, _, idx = model.encode(images_fn(b0, b1))
        hist += model.quantize._last_hist
        sums[0:1] += model.quantize._last_sse
        sums[1] += z.numel()
        out.append(idx)
    allreduce_usage(hist, sums, group)
    idx = torch.cat(out) if out else torch.empty(0, 0, dtype=torch.int64, device=dev)
    return idx, hist, global_loss(sums, model.quantize.beta)
