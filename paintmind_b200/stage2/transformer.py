"""CondTransformer — parameter container + CUDA forward with the reference's surface and
state_dict keys (stage2/transformer.py:28-93): token_proj, position_embedding, context_proj
(only when context_dim != dim), layers.layer{i}.{norm1,attn1,norm2,attn2,norm3,ffnet}, norm, to_logits."""
from __future__ import annotations

import torch
from torch import nn

from ..modules.attention import CrossAttention
from ..modules.mlp import SwiGLUFFNFused
from ..stage1.layers import _init_vit_weights


class Layer(nn.Module):
    """x = attn1(norm1(x)) + x; x = attn2(norm2(x), context) + x; x = ffnet(norm3(x)) + x  (transformer.py:44-49)."""

    def __init__(self, dim, dim_head, mlp_dim, num_head=8, dropout=0.0, dim_context=None):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = CrossAttention(query_dim=dim, heads=num_head, dim_head=dim_head, dropout=dropout)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=dim_context, heads=num_head, dim_head=dim_head, dropout=dropout)
        self.norm3 = nn.LayerNorm(dim)
        self.ffnet = SwiGLUFFNFused(in_features=dim, hidden_features=mlp_dim)


class CondTransformer(nn.Module):
    def __init__(self, in_dim, dim, len_seq, dim_head, mlp_dim, num_head=8, depth=6, dropout=0.1, context_dim=None,
                 num_classes=8192):
        super().__init__()
        self.in_dim, self.dim, self.len_seq, self.num_head, self.depth = in_dim, dim, len_seq, num_head, depth
        self.num_classes = num_classes
        self.token_proj = nn.Linear(in_dim, dim)
        self.position_embedding = nn.Parameter(torch.randn(1, len_seq, dim) * dim ** -0.5)
        self.context_proj = nn.Linear(context_dim, dim, bias=False) if context_dim != dim else nn.Identity()
        self.layers = nn.Sequential()
        for i in range(depth):
            self.layers.add_module("layer" + str(i), Layer(dim, dim_head, mlp_dim, num_head, dropout, dim))
        self.norm = nn.LayerNorm(dim)
        self.to_logits = nn.Linear(dim, num_classes)
        self.apply(_init_vit_weights)
        self._engine = None

    def engine(self):
        if self._engine is None:
            from ..engine import Stage2Engine
            object.__setattr__(self, "_engine", Stage2Engine(self))
        return self._engine

    @torch.no_grad()
    def forward(self, x, context=None):
        """x: tokens [B, N, in_dim]; context: [B, L, context_dim] or None -> fp32 logits [B, N, num_classes]."""
        return self.engine().forward(x, context)
