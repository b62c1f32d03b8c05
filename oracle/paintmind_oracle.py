"""CPU oracle for the PaintMind tokenizer hot path — TEST INFRASTRUCTURE ONLY.

A plain-numpy (fp32) restatement of the reference algorithm (Qiyuan-Ge/PaintMind, mounted snapshot
v0.0.0) for the path named by BASELINE.json: vit-s-vqgan encode -> quantize -> decode and the
MaskGIT stage-2 step.  Every function cites the reference file:line it follows.

  * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
    import this module.  The product path (paintmind_b200/) never does: it fails loudly when the
    CUDA library is missing.
  * PINNING: the reference ships no tests, golden vectors or KATs for this path,
    so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/make_golden.py imports
    /root/reference (read-only) in the build container, loads the same seeded state_dict into the
    reference modules and into this oracle, and commits the reference's outputs under
    tests/golden/.  tests/test_oracle_golden.py checks this file against those fixtures.
  * Third-party arithmetic the reference relies on (not under /root/reference): torch>=1.13
    (unpinned, setup.py:20; nn.Linear / LayerNorm / Conv2d / softmax / argmin / F.normalize),
    einops (unpinned).  Their published semantics are restated below in numpy.

Weights are passed as {state_dict key: np.ndarray(float32)} using the reference's key names
(SURVEY.md Appendix A).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------
# primitive ops (torch semantics)
# --------------------------------------------------------------------------------------------
def linear(x, w, b=None):
    """nn.Linear: y = x W^T + b."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y.astype(F32, copy=False)


def layer_norm(x, w, b, eps=1e-5):
    """nn.LayerNorm over the last dim, biased variance, eps=1e-5 (stage1/layers.py:49,51,89,128)."""
    mu = x.mean(axis=-1, keepdims=True, dtype=F32)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=F32)
    return (xc / np.sqrt(var + F32(eps)) * w + b).astype(F32, copy=False)


def softmax_lastdim(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m, dtype=F32)
    return e / e.sum(axis=-1, keepdims=True, dtype=F32)


def silu(x):
    return x / (F32(1.0) + np.exp(-x, dtype=F32))


def l2norm(t, eps=1e-12):
    """F.normalize(t, p=2, dim=-1): t / max(||t||_2, eps)  (stage1/quantize.py:5-6)."""
    n = np.sqrt((t * t).sum(axis=-1, keepdims=True, dtype=F32))
    return (t / np.maximum(n, F32(eps))).astype(F32, copy=False)


# --------------------------------------------------------------------------------------------
# modules/attention.py:43-59 (CrossAttention.forward) — identical math to the xformers class
# --------------------------------------------------------------------------------------------
def attention(x, sd, prefix, heads, context=None):
    ctx = x if context is None else context                        # attention.py:47
    q = linear(x, sd[prefix + "to_q.weight"])                      # :46 (no bias)
    k = linear(ctx, sd[prefix + "to_k.weight"])                    # :48
    v = linear(ctx, sd[prefix + "to_v.weight"])                    # :49
    B, N, inner = q.shape
    L = k.shape[1]
    d = inner // heads
    q = q.reshape(B, N, heads, d).transpose(0, 2, 1, 3)            # 'b n (h d) -> (b h) n d' :51
    k = k.reshape(B, L, heads, d).transpose(0, 2, 1, 3)
    v = v.reshape(B, L, heads, d).transpose(0, 2, 1, 3)
    q = q * F32(d ** -0.5)                                         # :52
    sim = q @ k.transpose(0, 1, 3, 2)                              # :54
    sim = softmax_lastdim(sim)                                     # :55
    out = sim @ v                                                  # :57
    out = out.transpose(0, 2, 1, 3).reshape(B, N, inner)           # :58
    return linear(out, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])  # :59


# --------------------------------------------------------------------------------------------
# modules/mlp.py:27-31 (SwiGLUFFN.forward)
# --------------------------------------------------------------------------------------------
def swiglu_ffn(x, sd, prefix):
    x12 = linear(x, sd[prefix + "w12.weight"], sd[prefix + "w12.bias"])   # mlp.py:28
    h = x12.shape[-1] // 2
    x1, x2 = x12[..., :h], x12[..., h:]                                   # chunk(2) :29
    hidden = silu(x1) * x2                                                # :30
    return linear(hidden, sd[prefix + "w3.weight"], sd[prefix + "w3.bias"])  # :31


def swiglu_hidden(mlp_dim):
    """modules/mlp.py:53."""
    return (int(mlp_dim * 2 / 3) + 7) // 8 * 8


# --------------------------------------------------------------------------------------------
# stage1/layers.py
# --------------------------------------------------------------------------------------------
def vit_layer(x, sd, prefix, heads):
    """Layer.forward (layers.py:54-58)."""
    x = attention(layer_norm(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"]), sd, prefix + "attn1.", heads) + x
    x = swiglu_ffn(layer_norm(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"]), sd, prefix + "ffnet.") + x
    return x


def patch_embed(img, w):
    """Conv2d(C, dim, k=P, s=P, bias=False) + 'b c h w -> b (h w) c' (layers.py:81-84)."""
    B, C, H, W = img.shape
    dim, _, P, _ = w.shape
    gh, gw = H // P, W // P
    cols = img.reshape(B, C, gh, P, gw, P).transpose(0, 2, 4, 1, 3, 5).reshape(B, gh * gw, C * P * P)
    return (cols @ w.reshape(dim, C * P * P).T).astype(F32, copy=False)


def encoder_forward(img, sd, cfg, prefix="encoder."):
    """Encoder.forward (layers.py:106-112)."""
    x = patch_embed(img, sd[prefix + "to_patch_embedding.0.weight"])
    x = x + sd[prefix + "position_embedding"]
    x = layer_norm(x, sd[prefix + "norm_pre.weight"], sd[prefix + "norm_pre.bias"])
    for i in range(cfg["depth"]):
        x = vit_layer(x, sd, f"{prefix}transformer.layers.{i}.", cfg["num_head"])
    return x


def decoder_forward(x, sd, cfg, prefix="decoder."):
    """Decoder.forward (layers.py:145-152)."""
    x = x + sd[prefix + "position_embedding"]
    for i in range(cfg["depth"]):
        x = vit_layer(x, sd, f"{prefix}transformer.layers.{i}.", cfg["num_head"])
    x = layer_norm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    x = linear(x, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])
    B = x.shape[0]
    P = cfg["patch_size"]
    g = cfg["image_size"] // P
    C = x.shape[-1] // (P * P)
    # 'b (h w) (p1 p2 c) -> b c (h p1) (w p2)'  (layers.py:150)
    return x.reshape(B, g, g, P, P, C).transpose(0, 5, 1, 3, 2, 4).reshape(B, C, g * P, g * P)


# --------------------------------------------------------------------------------------------
# stage1/quantize.py
# --------------------------------------------------------------------------------------------
def vq_distances(z, codebook):
    """quantize.py:19-26: distances between normalised latents and the normalised codebook."""
    zn = l2norm(z)
    zf = zn.reshape(-1, codebook.shape[1])
    en = l2norm(codebook)
    d = (zf * zf).sum(axis=1, keepdims=True, dtype=F32) + (en * en).sum(axis=1, dtype=F32) - F32(2.0) * (zf @ en.T)
    return zn, zf, en, d.astype(F32, copy=False)


def vq_forward(z, codebook, beta=0.25):
    """VectorQuantizer.forward (quantize.py:18-38) -> (z_q, loss, indices[int64])."""
    zn, zf, en, d = vq_distances(z, codebook)
    idx = np.argmin(d, axis=1).astype(np.int64).reshape(zn.shape[:-1])     # first minimum, :28
    z_q = l2norm(codebook[idx])                                             # :29-30
    diff = z_q - zn
    mse = (diff * diff).mean(dtype=F32)
    loss = F32(beta) * mse + mse                                            # :33
    z_q = zn + (z_q - zn)                                                   # :36 (forward value)
    return z_q.astype(F32, copy=False), F32(loss), idx


def vq_top2_gap(z, codebook):
    """Per-latent gap between the two smallest distances (parity tolerance)."""
    _, _, _, d = vq_distances(z, codebook)
    part = np.partition(d, 1, axis=1)[:, :2]
    return (part[:, 1] - part[:, 0]).astype(F32)


def vq_decode_from_indice(idx, codebook):
    """quantize.py:40-44."""
    return l2norm(codebook[idx])


# --------------------------------------------------------------------------------------------
# stage1/vqmodel.py
# --------------------------------------------------------------------------------------------
def vqmodel_encode(img, sd, cfg):
    """VQModel.encode (vqmodel.py:21-25)."""
    x = encoder_forward(img, sd, cfg["enc"])
    x = linear(x, sd["prev_quant.weight"], sd["prev_quant.bias"])
    return vq_forward(x, sd["quantize.embedding.weight"], cfg["beta"])


def vqmodel_latent(img, sd, cfg):
    """encoder + prev_quant only (the un-normalised 32-d latent fed to the quantizer)."""
    x = encoder_forward(img, sd, cfg["enc"])
    return linear(x, sd["prev_quant.weight"], sd["prev_quant.bias"])


def vqmodel_decode(z, sd, cfg, clamp=True):
    """VQModel.decode (vqmodel.py:27-30)."""
    x = linear(z, sd["post_quant.weight"], sd["post_quant.bias"])
    x = decoder_forward(x, sd, cfg["dec"])
    return np.clip(x, -1.0, 1.0) if clamp else x


def vqmodel_decode_from_indice(idx, sd, cfg):
    """vqmodel.py:38-41."""
    return vqmodel_decode(vq_decode_from_indice(idx, sd["quantize.embedding.weight"]), sd, cfg)


# --------------------------------------------------------------------------------------------
# stage2/transformer.py
# --------------------------------------------------------------------------------------------
def cond_layer(x, context, sd, prefix, heads):
    """stage2 Layer.forward (transformer.py:44-49)."""
    x = attention(layer_norm(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"]), sd, prefix + "attn1.", heads) + x
    x = attention(layer_norm(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"]), sd, prefix + "attn2.", heads, context) + x
    x = swiglu_ffn(layer_norm(x, sd[prefix + "norm3.weight"], sd[prefix + "norm3.bias"]), sd, prefix + "ffnet.") + x
    return x


def cond_transformer_forward(tokens, context, sd, cfg, prefix="transformer."):
    """CondTransformer.forward (transformer.py:80-93)."""
    x = linear(tokens, sd[prefix + "token_proj.weight"], sd[prefix + "token_proj.bias"])
    x = x + sd[prefix + "position_embedding"]
    if context is not None and (prefix + "context_proj.weight") in sd:
        context = linear(context, sd[prefix + "context_proj.weight"])     # :58 (Identity when dims match)
    for i in range(cfg["depth"]):
        x = cond_layer(x, context, sd, f"{prefix}layers.layer{i}.", cfg["num_head"])
    x = layer_norm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    return linear(x, sd[prefix + "to_logits.weight"], sd[prefix + "to_logits.bias"])


# --------------------------------------------------------------------------------------------
# generate.py (MaskGIT sampling half)
# --------------------------------------------------------------------------------------------
def mask_schedule(ratio):
    """generate.py:25-26 (numpy float64)."""
    return np.cos(math.pi / 2.0 * ratio)


def ids2tokens(ids, codebook_raw, mask_token):
    """generate.py:148-157 — RAW (un-normalised) codebook + mask token."""
    table = np.concatenate([codebook_raw, mask_token], axis=0)
    return table[ids]


def top_k_filter(logits, k):
    """generate.py:33-37: keep the k largest logits per token, -inf elsewhere."""
    idx = np.argpartition(-logits, k - 1, axis=-1)[..., :k]
    out = np.full_like(logits, -np.inf)
    np.put_along_axis(out, idx, np.take_along_axis(logits, idx, axis=-1), axis=-1)
    return out


def gumbel_from_uniform(u):
    """generate.py:29-30,40-42: -log(-log(u)) with log(t) = log(clamp(t, 1e-20))."""
    inner = -np.log(np.maximum(u, F32(1e-20)), dtype=F32)
    return (-np.log(np.maximum(inner, F32(1e-20)), dtype=F32)).astype(F32)


def gumbel_sample(filtered, temperature, uniform):
    """generate.py:45-46 with the uniform noise injected by the caller."""
    return np.argmax(filtered / F32(max(temperature, 1e-10)) + gumbel_from_uniform(uniform), axis=-1).astype(np.int64)


def remask(ids, pred_ids, logits, mask_ratio, mask_id, num_tokens):
    """generate.py:166-179: fill masked positions, score = 1 - p(pred), re-mask the k least confident.

    Returns (new_ids, scores, k).  Ties at the k-th score are implementation-defined in the
    reference (torch.topk); callers compare sets modulo ties.
    """
    is_mask = ids == mask_id
    ids = np.where(is_mask, pred_ids, ids)
    probs = softmax_lastdim(logits)
    scores = F32(1.0) - np.take_along_axis(probs, pred_ids[..., None], axis=-1)[..., 0]
    scores = np.where(is_mask, scores, F32(-1e5)).astype(F32)
    k = max(int(mask_ratio * num_tokens), 1)
    order = np.argsort(-scores, axis=-1, kind="stable")[:, :k]
    out = ids.copy()
    np.put_along_axis(out, order, mask_id, axis=-1)
    return out, scores, k


def sample_step(ids, mask_ratio, text, topk, temperature, uniform, sd, cfg2, cfg1, decode=True):
    """Pipeline.sample (generate.py:159-181) with injected uniform noise."""
    mask_id = cfg1["n_embed"]
    tokens = ids2tokens(ids, sd["vqgan.quantize.embedding.weight"], sd["mask_token"])
    logits = cond_transformer_forward(tokens, text, sd, cfg2)
    pred_ids = gumbel_sample(top_k_filter(logits, topk), temperature, uniform)
    img = None
    if decode:
        sd1 = {k[len("vqgan."):]: v for k, v in sd.items() if k.startswith("vqgan.")}
        img = vqmodel_decode_from_indice(pred_ids, sd1, cfg1)
    num_tokens = (cfg1["enc"]["image_size"] // cfg1["enc"]["patch_size"]) ** 2
    new_ids, scores, k = remask(ids, pred_ids, logits, mask_ratio, mask_id, num_tokens)
    return new_ids, img, pred_ids, logits, scores, k


def generate_schedule(timesteps, temperature, num_tokens):
    """Host scalars of Pipeline.generate (generate.py:190-193): per step (mask_ratio, k, cur_temp)."""
    out = []
    for step in range(timesteps):
        r = mask_schedule((step + 1) / timesteps)
        out.append((float(r), max(int(r * num_tokens), 1), temperature * (1 - step / timesteps)))
    return out


# --------------------------------------------------------------------------------------------
# generate.py (stage-2 TRAINING forward: random_masking / loss / forward, :78-146)
# --------------------------------------------------------------------------------------------
def random_masking(x, mask_ratio, mask_token, noise):
    """Pipeline.random_masking (generate.py:78-108) with the uniform noise injected by the caller.
    Returns (x with masked rows replaced by mask_token, mask [N, L] fp32, 1 = masked).  Restated literally:
    argsort -> keep the first len_keep -> append mask tokens -> un-shuffle.  (numpy's stable argsort keeps the
    lower index first on exact noise ties; torch.argsort leaves that unspecified.)"""
    N, L, D = x.shape
    len_mask = max(int(L * mask_ratio), 1)                                  # :86
    len_keep = L - len_mask                                                 # :87
    ids_shuffle = np.argsort(noise, axis=1, kind="stable")                  # :92
    ids_restore = np.argsort(ids_shuffle, axis=1, kind="stable")            # :93
    ids_keep = ids_shuffle[:, :len_keep]                                    # :95
    xk = np.take_along_axis(x, ids_keep[..., None], axis=1)                 # :96
    mt = np.broadcast_to(mask_token.reshape(1, 1, D), (N, L - len_keep, D)) # :97
    xc = np.concatenate([xk, mt], axis=1)                                   # :98
    xo = np.take_along_axis(xc, ids_restore[..., None], axis=1)             # :100
    mask = np.ones((N, L), dtype=F32)                                       # :103
    mask[:, :len_keep] = 0                                                  # :104
    mask = np.take_along_axis(mask, ids_restore, axis=1)                    # :106
    return xo.astype(F32, copy=False), mask


def ce_label_smooth_rows(logits2d, label, eps=0.1):
    """F.cross_entropy(logit, label, label_smoothing=eps, reduction='none') (generate.py:121):
    (1 - eps) * nll(label) + eps * mean_c(-log p_c)."""
    m = logits2d.max(axis=-1, keepdims=True)
    lse = (m[:, 0] + np.log(np.exp(logits2d - m, dtype=F32).sum(axis=-1, dtype=F32))).astype(F32)
    xy = np.take_along_axis(logits2d, label[:, None], axis=-1)[:, 0]
    return (lse - F32(1.0 - eps) * xy - F32(eps) * logits2d.mean(axis=-1, dtype=F32)).astype(F32)


def masked_ce_loss(logits, label, masks, eps=0.1):
    """Pipeline.loss (generate.py:110-123)."""
    V = logits.shape[-1]
    rows = ce_label_smooth_rows(logits.reshape(-1, V), label.reshape(-1), eps)
    mk = masks.reshape(-1).astype(F32)
    return F32((rows * mk).sum(dtype=np.float64) / mk.sum(dtype=np.float64))


def pipeline_train_forward(img, text, mask_ratio, noise, sd, cfg2, cfg1):
    """Pipeline.forward (generate.py:136-146): to_latent -> random_masking -> tokens2logits -> loss."""
    sd1 = {k[len("vqgan."):]: v for k, v in sd.items() if k.startswith("vqgan.")}
    z_q, _, ids = vqmodel_encode(img, sd1, cfg1)                                    # :138 (to_latent, :125-131)
    xm, mask = random_masking(z_q, mask_ratio, sd["mask_token"], noise)             # :140
    logits = cond_transformer_forward(xm, text, sd, cfg2)                           # :142
    return masked_ce_loss(logits, ids, mask), ids, mask, logits                     # :144
