"""Seeded synthetic weights and inputs (there is no network for checkpoints or datasets).

The SAME state_dict is loaded into the reference modules (when generating golden fixtures), into
the numpy oracle and into the CUDA modules, so parity never depends on RNG call order inside a
constructor.  Distributions follow the reference initialisers except that
LayerNorm affine parameters and biases are perturbed away from (1, 0) so that every term of the
arithmetic is exercised.
"""
from __future__ import annotations

import torch

from ..modules.mlp import swiglu_hidden


def _xavier(g, out_f, in_f):
    bound = (6.0 / (in_f + out_f)) ** 0.5
    return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound


def _ln(g, sd, prefix, dim):
    sd[prefix + "weight"] = 1.0 + 0.1 * torch.randn(dim, generator=g)
    sd[prefix + "bias"] = 0.05 * torch.randn(dim, generator=g)


def _attn(g, sd, prefix, dim, inner, ctx_dim=None):
    ctx_dim = dim if ctx_dim is None else ctx_dim
    sd[prefix + "to_q.weight"] = _xavier(g, inner, dim)
    sd[prefix + "to_k.weight"] = _xavier(g, inner, ctx_dim)
    sd[prefix + "to_v.weight"] = _xavier(g, inner, ctx_dim)
    sd[prefix + "to_out.0.weight"] = _xavier(g, dim, inner)
    sd[prefix + "to_out.0.bias"] = 0.02 * torch.randn(dim, generator=g)


def _ffn(g, sd, prefix, dim, mlp_dim):
    h = swiglu_hidden(mlp_dim)
    sd[prefix + "w12.weight"] = _xavier(g, 2 * h, dim)
    sd[prefix + "w12.bias"] = 0.02 * torch.randn(2 * h, generator=g)
    sd[prefix + "w3.weight"] = _xavier(g, dim, h)
    sd[prefix + "w3.bias"] = 0.02 * torch.randn(dim, generator=g)


def _vit_layers(g, sd, prefix, cfg):
    dim, inner = cfg["dim"], cfg["num_head"] * cfg["dim_head"]
    for i in range(cfg["depth"]):
        p = f"{prefix}transformer.layers.{i}."
        _ln(g, sd, p + "norm1.", dim)
        _attn(g, sd, p + "attn1.", dim, inner)
        _ln(g, sd, p + "norm2.", dim)
        _ffn(g, sd, p + "ffnet.", dim, cfg["mlp_dim"])


def make_vqgan_state_dict(cfg: dict, seed: int = 0) -> dict:
    """fp32 CPU state_dict with the reference VQModel's 222 keys (for vit-s-vqgan)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    enc, dec = cfg["enc"], cfg["dec"]
    n_tok = (enc["image_size"] // enc["patch_size"]) ** 2
    fan_in = enc["in_channels"] * enc["patch_size"] ** 2
    sd["encoder.position_embedding"] = torch.randn(1, n_tok, enc["dim"], generator=g) * enc["dim"] ** -0.5
    sd["encoder.to_patch_embedding.0.weight"] = (torch.rand(enc["dim"], enc["in_channels"], enc["patch_size"], enc["patch_size"], generator=g) * 2 - 1) * fan_in ** -0.5
    _ln(g, sd, "encoder.norm_pre.", enc["dim"])
    _vit_layers(g, sd, "encoder.", enc)
    sd["decoder.position_embedding"] = torch.randn(1, n_tok, dec["dim"], generator=g) * dec["dim"] ** -0.5
    _vit_layers(g, sd, "decoder.", dec)
    _ln(g, sd, "decoder.norm.", dec["dim"])
    n_out = dec["out_channels"] * dec["patch_size"] ** 2
    sd["decoder.proj.weight"] = _xavier(g, n_out, dec["dim"])
    sd["decoder.proj.bias"] = 0.02 * torch.randn(n_out, generator=g)
    sd["quantize.embedding.weight"] = torch.randn(cfg["n_embed"], cfg["embed_dim"], generator=g)
    sd["prev_quant.weight"] = (torch.rand(cfg["embed_dim"], enc["dim"], generator=g) * 2 - 1) * enc["dim"] ** -0.5
    sd["prev_quant.bias"] = (torch.rand(cfg["embed_dim"], generator=g) * 2 - 1) * enc["dim"] ** -0.5
    sd["post_quant.weight"] = (torch.rand(dec["dim"], cfg["embed_dim"], generator=g) * 2 - 1) * cfg["embed_dim"] ** -0.5
    sd["post_quant.bias"] = (torch.rand(dec["dim"], generator=g) * 2 - 1) * cfg["embed_dim"] ** -0.5
    return sd


def make_stage2_state_dict(cfg2: dict, cfg1: dict, seed: int = 1, context_dim: int = 1024) -> dict:
    """fp32 CPU state_dict for CondTransformer + mask_token (keys as in Pipeline, minus vqgan./text_model.)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    dim, inner = cfg2["dim"], cfg2["num_head"] * cfg2["dim_head"]
    n_tok = (cfg1["enc"]["image_size"] // cfg1["enc"]["patch_size"]) ** 2
    sd["mask_token"] = 0.02 * torch.randn(1, cfg1["embed_dim"], generator=g)
    p = "transformer."
    sd[p + "position_embedding"] = torch.randn(1, n_tok, dim, generator=g) * dim ** -0.5
    sd[p + "token_proj.weight"] = _xavier(g, dim, cfg1["embed_dim"])
    sd[p + "token_proj.bias"] = 0.02 * torch.randn(dim, generator=g)
    if context_dim != dim:
        sd[p + "context_proj.weight"] = _xavier(g, dim, context_dim)
    for i in range(cfg2["depth"]):
        lp = f"{p}layers.layer{i}."
        _ln(g, sd, lp + "norm1.", dim)
        _attn(g, sd, lp + "attn1.", dim, inner)
        _ln(g, sd, lp + "norm2.", dim)
        _attn(g, sd, lp + "attn2.", dim, inner, dim)
        _ln(g, sd, lp + "norm3.", dim)
        _ffn(g, sd, lp + "ffnet.", dim, cfg2["mlp_dim"])
    _ln(g, sd, p + "norm.", dim)
    sd[p + "to_logits.weight"] = _xavier(g, cfg1["n_embed"], dim)
    sd[p + "to_logits.bias"] = 0.02 * torch.randn(cfg1["n_embed"], generator=g)
    return sd


def make_images(batch: int, image_size: int = 256, seed: int = 0, channels: int = 3) -> torch.Tensor:
    """fp32 NCHW in [-1, 1] (the range stage1_transform produces, utils/transform.py:17-18)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, channels, image_size, image_size, generator=g) * 2 - 1
