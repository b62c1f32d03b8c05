"""Attention parameter container with the reference's state_dict keys.

Reference: modules/attention.py:25-59 (CrossAttention) and :62-108 (MemoryEfficientCrossAttention)
hold identical parameters: to_q / to_k / to_v (no bias) and to_out = Sequential(Linear, Dropout).
Here the math lives in csrc/pm_attn.cu + the projection GEMMs; this module only owns parameters and
offers a stand-alone forward that runs those kernels (used by the unit parity tests).
"""
from __future__ import annotations

import torch
from torch import nn


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))

    @torch.no_grad()
    def forward(self, x, context=None):
        from ..engine import standalone_attention
        return standalone_attention(self, x, context)


# Same parameters, same kernels: the "xformers" flavour of the reference is just another name here.
MemoryEfficientCrossAttention = CrossAttention
XFORMERS_IS_AVAILBLE = False  # (sic) name kept from modules/attention.py:7-12; no xformers dependency
