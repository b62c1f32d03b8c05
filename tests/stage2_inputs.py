"""Seeded inputs shared by the stage-2 fixture generator and the tests (must stay in sync)."""
import torch

TINY2 = dict(dim=128, dim_head=64, mlp_dim=256, num_head=2, depth=2, dropout=0.1)


def tiny_inputs(ctx_dim, seed=11):
    g = torch.Generator().manual_seed(seed)
    tokens = torch.randn(2, 64, 32, generator=g)
    context = torch.randn(2, 77, ctx_dim, generator=g)
    return tokens, context


def full_step_inputs(seed=21, B=1, N=1024, V=8192):
    g = torch.Generator().manual_seed(seed)
    text = torch.randn(B, 77, 1024, generator=g)
    ids = torch.full((B, N), V, dtype=torch.long)
    keep = torch.rand(B, N, generator=g) < 0.3
    ids[keep] = torch.randint(0, V, (int(keep.sum()),), generator=g)
    u = torch.rand(B, N, V, generator=g)
    return text, ids, u
