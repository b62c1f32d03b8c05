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


def allreduce_gradients(module, group=None, average: bool = True, bucket: torch.Tensor | None = None):
    """Data-parallel generator training (the reference wraps its trainer in accelerate / DDP, utils/trainer.py:74-90,216):
    every rank runs the training step on its own batch shard with replicated weights; the only exchange is ONE
    all-reduce of the parameter gradients, flattened into a single fp32 bucket (208 MB for vit-s-vqgan: one NCCL call
    over NVLink instead of 222 small ones), averaged over the ranks and scattered back into `.grad`.
    Returns the bucket (pass it back in to reuse the allocation)."""
    params = [p for p in module.parameters() if p.grad is not None]
    if not params:
        return bucket
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return bucket
    n = sum(p.grad.numel() for p in params)
    dev = params[0].grad.device
    if bucket is None or bucket.numel() < n or bucket.device != dev:
        bucket = torch.empty(n, device=dev, dtype=torch.float32)
    flat = bucket[:n]
    views, off = [], 0
    for p in params:
        k = p.grad.numel()
        views.append(flat[off:off + k].view_as(p.grad))
        off += k
    torch._foreach_copy_(views, [p.grad for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.mul_(1.0 / world)
    torch._foreach_copy_([p.grad for p in params], views)
    return bucket
