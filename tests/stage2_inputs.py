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


# ---- stage-2 TRAINING forward (generate.py:78-146) ----
def masking_inputs(seed=31, B=2, N=1024, D=32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, D, generator=g)
    noise = torch.rand(B, N, generator=g)
    noise[0, 5] = noise[0, 900]                      # one exact tie in the noise (tie rule: lower index kept first)
    return x, noise


def loss_inputs(seed=32, B=2, L=96, V=8192):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, L, V, generator=g) * 3.0
    logits[0, 0, 17] = 40.0                          # a confident row
    label = torch.randint(0, V, (B, L), generator=g)
    label[0, 0] = 17
    masks = (torch.rand(B, L, generator=g) < 0.7).float()
    return logits, label, masks


def train_forward_inputs(seed=33, B=1):
    from paintmind_b200.utils import synthetic
    g = torch.Generator().manual_seed(seed)
    img = synthetic.make_images(B, 256, seed=seed + 100)
    noise = torch.rand(B, 1024, generator=g)
    return img, noise
