"""SwiGLU feed-forward parameter container (reference modules/mlp.py:13-59).

w12: Linear(in, 2*hidden) — first half gate, second half value (mlp.py:28-30); w3: Linear(hidden, out).
SwiGLUFFNFused rescales hidden to 2/3 and rounds up to a multiple of 8 (mlp.py:51-53).
The math runs in csrc/pm_gemm.cu (w12 GEMM with the silu-gate epilogue, w3 GEMM with residual).
"""
from __future__ import annotations

import torch
from torch import nn


def swiglu_hidden(hidden_features: int) -> int:
    return (int(hidden_features * 2 / 3) + 7) // 8 * 8


class SwiGLUFFN(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, bias=True):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.w12 = nn.Linear(in_features, 2 * hidden_features, bias=bias)
        self.w3 = nn.Linear(hidden_features, out_features, bias=bias)

    @torch.no_grad()
    def forward(self, x):
        from ..engine import standalone_swiglu
        return standalone_swiglu(self, x)


class SwiGLUFFNFused(SwiGLUFFN):
    def __init__(self, in_features, hidden_features=None, out_features=None, bias=True):
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        super().__init__(in_features, swiglu_hidden(hidden_features), out_features, bias)
