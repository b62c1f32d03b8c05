"""ViT encoder / decoder parameter containers with the reference's state_dict keys
(stage1/layers.py:40-152).  forward() of each module runs the CUDA engine (engine.py)."""
from __future__ import annotations

import torch
from torch import nn

from ..modules.attention import CrossAttention
from ..modules.mlp import SwiGLUFFNFused


def _init_vit_weights(m):
    # reference Encoder/Decoder._init_weights (layers.py:97-104, 136-143)
    if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


class Layer(nn.Module):
    """Pre-LN block: x = attn1(norm1(x)) + x; x = ffnet(norm2(x)) + x  (layers.py:40-58)."""

    def __init__(self, dim, dim_head, mlp_dim, num_head=8, dropout=0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = CrossAttention(query_dim=dim, heads=num_head, dim_head=dim_head, dropout=dropout)
        self.norm2 = nn.LayerNorm(dim)
        self.ffnet = SwiGLUFFNFused(in_features=dim, hidden_features=mlp_dim)


class Transformer(nn.Module):
    def __init__(self, dim, depth, num_head, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.layers = nn.Sequential(*[Layer(dim, dim_head, mlp_dim, num_head, dropout) for _ in range(depth)])


class _TokensFromMap(nn.Module):
    """Placeholder for einops' Rearrange('b c h w -> b (h w) c') at index 1 of to_patch_embedding
    (layers.py:83); parameter-free, keeps the Sequential indices (and state_dict keys) identical."""

    def forward(self, x):
        return x.flatten(2).transpose(1, 2)


class Encoder(nn.Module):
    def __init__(self, image_size, patch_size, dim, depth, num_head, mlp_dim, in_channels=3, out_channels=3,
                 dim_head=64, dropout=0.0):
        super().__init__()
        assert image_size % patch_size == 0, "Image dimensions must be divisible by the patch size."
        self.image_size, self.patch_size = image_size, patch_size
        self.dim, self.depth, self.num_head, self.dim_head = dim, depth, num_head, dim_head
        self.in_channels = in_channels
        self.to_patch_embedding = nn.Sequential(
            nn.Conv2d(in_channels, dim, kernel_size=patch_size, stride=patch_size, bias=False),
            _TokensFromMap(),
        )
        num_patches = (image_size // patch_size) ** 2
        self.position_embedding = nn.Parameter(torch.randn(1, num_patches, dim) * dim ** -0.5)
        self.norm_pre = nn.LayerNorm(dim)
        self.transformer = Transformer(dim, depth, num_head, dim_head, mlp_dim, dropout)
        self.apply(_init_vit_weights)

    @torch.no_grad()
    def forward(self, x):
        from ..engine import engine_for
        return engine_for(self).run_encoder(x).float()


class Decoder(nn.Module):
    def __init__(self, image_size, patch_size, dim, depth, num_head, mlp_dim, in_channels=3, out_channels=3,
                 dim_head=64, dropout=0.0):
        super().__init__()
        assert image_size % patch_size == 0, "Image dimensions must be divisible by the patch size."
        self.image_size, self.patch_size = image_size, patch_size
        self.dim, self.depth, self.num_head, self.dim_head = dim, depth, num_head, dim_head
        self.out_channels = out_channels
        num_patches = (image_size // patch_size) ** 2
        self.position_embedding = nn.Parameter(torch.randn(1, num_patches, dim) * dim ** -0.5)
        self.transformer = Transformer(dim, depth, num_head, dim_head, mlp_dim, dropout)
        self.norm = nn.LayerNorm(dim)
        self.proj = nn.Linear(dim, out_channels * patch_size * patch_size, bias=True)
        self.apply(_init_vit_weights)

    @torch.no_grad()
    def forward(self, x):
        from ..engine import engine_for
        return engine_for(self).run_decoder_tokens(x)
