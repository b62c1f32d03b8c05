"""CPU oracle, torch flavour — TEST / BASELINE INFRASTRUCTURE ONLY (never on the product path).

Same restatement as oracle/paintmind_oracle.py (every function cites the reference file:line it
follows), but expressed with torch's CPU operators in fp32, i.e. the very ATen kernels
(`addmm`, `bmm`, `_softmax`, `native_layer_norm`, `silu`, multithreaded through OpenMP/MKL) that the
pure-Python reference dispatches to when it runs on the host.  Its purpose is the *timed* CPU arm
of bench.py (`cpu_baseline`, `--impl reference`): the numpy port is 2-3x slower than the reference
really is on the same cores (single-threaded exp / softmax), which would flatter the GPU/CPU ratio.
Functional code only: no nn.Module, nothing imported from /root/reference (which does not exist on
the GPU box).  Pinned by tests/test_oracle_golden.py against the same reference-generated fixtures
as the numpy oracle.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def attention(x, sd, prefix, heads, context=None):
    """CrossAttention.forward (modules/attention.py:43-59)."""
    ctx = x if context is None else context                                    # :47
    q = F.linear(x, sd[prefix + "to_q.weight"])                                # :46
    k = F.linear(ctx, sd[prefix + "to_k.weight"])                              # :48
    v = F.linear(ctx, sd[prefix + "to_v.weight"])                              # :49
    B, N, inner = q.shape
    L = k.shape[1]
    d = inner // heads
    # 'b n (h d) -> (b h) n d'  (:51)
    q = q.view(B, N, heads, d).permute(0, 2, 1, 3).reshape(B * heads, N, d)
    k = k.view(B, L, heads, d).permute(0, 2, 1, 3).reshape(B * heads, L, d)
    v = v.view(B, L, heads, d).permute(0, 2, 1, 3).reshape(B * heads, L, d)
    q = q * d ** -0.5                                                          # :52
    sim = torch.bmm(q, k.transpose(1, 2))                                      # :54 (einsum 'b i d, b j d -> b i j')
    sim = sim.softmax(dim=-1)                                                  # :55
    out = torch.bmm(sim, v)                                                    # :57
    out = out.view(B, heads, N, d).permute(0, 2, 1, 3).reshape(B, N, inner)    # :58
    return F.linear(out, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])   # :59


def swiglu_ffn(x, sd, prefix):
    """SwiGLUFFN.forward (modules/mlp.py:27-31)."""
    x12 = F.linear(x, sd[prefix + "w12.weight"], sd[prefix + "w12.bias"])
    x1, x2 = x12.chunk(2, dim=-1)
    return F.linear(F.silu(x1) * x2, sd[prefix + "w3.weight"], sd[prefix + "w3.bias"])


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def vit_layer(x, sd, prefix, heads):
    """Layer.forward (stage1/layers.py:54-58)."""
    x = attention(_ln(x, sd, prefix + "norm1"), sd, prefix + "attn1.", heads) + x
    x = swiglu_ffn(_ln(x, sd, prefix + "norm2"), sd, prefix + "ffnet.") + x
    return x


def encoder_forward(img, sd, cfg, prefix="encoder."):
    """Encoder.forward (stage1/layers.py:106-112); patch embedding = Conv2d(k=P, s=P, bias=False) (:81-84)."""
    P = cfg["patch_size"]
    x = F.conv2d(img, sd[prefix + "to_patch_embedding.0.weight"], stride=P)
    x = x.flatten(2).transpose(1, 2)                                           # 'b c h w -> b (h w) c'
    x = x + sd[prefix + "position_embedding"]
    x = _ln(x, sd, prefix + "norm_pre")
    for i in range(cfg["depth"]):
        x = vit_layer(x, sd, f"{prefix}transformer.layers.{i}.", cfg["num_head"])
    return x


def decoder_forward(x, sd, cfg, prefix="decoder."):
    """Decoder.forward (stage1/layers.py:145-152)."""
    x = x + sd[prefix + "position_embedding"]
    for i in range(cfg["depth"]):
        x = vit_layer(x, sd, f"{prefix}transformer.layers.{i}.", cfg["num_head"])
    x = _ln(x, sd, prefix + "norm")
    x = F.linear(x, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])
    B = x.shape[0]
    P = cfg["patch_size"]
    g = cfg["image_size"] // P
    C = x.shape[-1] // (P * P)
    # 'b (h w) (p1 p2 c) -> b c (h p1) (w p2)'  (:150)
    return x.view(B, g, g, P, P, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, g * P, g * P)


def vq_forward(z, codebook, beta=0.25):
    """VectorQuantizer.forward (stage1/quantize.py:18-38) -> (z_q, loss, indices int64)."""
    zn = F.normalize(z, p=2, dim=-1)                                           # :19
    zf = zn.reshape(-1, codebook.shape[1])                                     # :20
    en = F.normalize(codebook, p=2, dim=-1)                                    # :21
    d = zf.pow(2).sum(dim=1, keepdim=True) + en.pow(2).sum(dim=1) - 2 * (zf @ en.t())    # :23-26
    idx = torch.argmin(d, dim=1).view(zn.shape[:-1])                           # :28
    z_q = F.normalize(codebook[idx], p=2, dim=-1)                              # :29-30
    mse = torch.mean((z_q - zn) ** 2)
    loss = beta * mse + mse                                                    # :33
    z_q = zn + (z_q - zn)                                                      # :36
    return z_q, loss, idx


def vqmodel_encode(img, sd, cfg):
    """VQModel.encode (stage1/vqmodel.py:21-25)."""
    x = encoder_forward(img, sd, cfg["enc"])
    x = F.linear(x, sd["prev_quant.weight"], sd["prev_quant.bias"])
    return vq_forward(x, sd["quantize.embedding.weight"], cfg["beta"])


def vqmodel_latent(img, sd, cfg):
    x = encoder_forward(img, sd, cfg["enc"])
    return F.linear(x, sd["prev_quant.weight"], sd["prev_quant.bias"])


def vqmodel_decode(z, sd, cfg, clamp=True):
    """VQModel.decode (stage1/vqmodel.py:27-30)."""
    x = F.linear(z, sd["post_quant.weight"], sd["post_quant.bias"])
    x = decoder_forward(x, sd, cfg["dec"])
    return x.clamp(-1.0, 1.0) if clamp else x


# ------------------------------------------------------------------------------------------------
# training forward (what VQGANTrainer differentiates, utils/trainer.py:205-217): same functions, run under
# autograd on a state dict whose tensors require grad; only the quantizer needs its detach structure spelled out.
# ------------------------------------------------------------------------------------------------
def vq_forward_train(z, codebook, beta=0.25, idx=None):
    """VectorQuantizer.forward with the reference's stop-gradients (stage1/quantize.py:18-38).  `idx` overrides the
    argmin (tests pass the indices the CUDA path chose, so that near-tie flips do not enter a gradient comparison)."""
    zn = F.normalize(z, p=2, dim=-1)                                           # :19
    if idx is None:
        zf = zn.reshape(-1, codebook.shape[1])
        en = F.normalize(codebook, p=2, dim=-1)                                # :21
        d = zf.pow(2).sum(dim=1, keepdim=True) + en.pow(2).sum(dim=1) - 2 * (zf @ en.t())
        idx = torch.argmin(d, dim=1).view(zn.shape[:-1])                       # :28
    z_q = F.normalize(codebook[idx], p=2, dim=-1)                              # :29-30
    loss = beta * torch.mean((z_q.detach() - zn) ** 2) + torch.mean((z_q - zn.detach()) ** 2)   # :33
    z_q = zn + (z_q - zn).detach()                                             # :36
    return z_q, loss, idx


def vqmodel_forward_train(img, sd, cfg, idx=None):
    """VQModel.forward (stage1/vqmodel.py:32-36) -> (rec, codebook loss, indices), differentiable."""
    x = encoder_forward(img, sd, cfg["enc"])
    x = F.linear(x, sd["prev_quant.weight"], sd["prev_quant.bias"])
    z_q, loss, idx = vq_forward_train(x, sd["quantize.embedding.weight"], cfg["beta"], idx)
    return vqmodel_decode(z_q, sd, cfg), loss, idx
