"""Generator training step of the tokenizer: VQModel.forward WITH autograd.

The reference has no backward code of its own: `VQGANTrainer.train` (utils/trainer.py:205-225) runs
`rec, codebook_loss = self.vqvae(img)` under autograd and calls `accelerator.backward(loss)`; torch differentiates
stage1/vqmodel.py:21-36, stage1/layers.py:54-58,106-112,145-152, modules/attention.py:43-59, modules/mlp.py:27-31 and
stage1/quantize.py:18-38.  Here the same step is ONE `torch.autograd.Function` whose forward and backward sequence the
hand-written sm_100a kernels of libpaintmind_b200 (ops.py); torch only routes the incoming `d rec`, `d loss` and hands the
222 parameter gradients to its optimiser.  LPIPS / discriminator / optimiser stay PyTorch (out of scope).

Memory plan (M = B * 1024 tokens): the forward keeps ONE bf16 [M, D] checkpoint per transformer block (the block input)
plus the few tensors at the path's joints (patch rows, pre-norm tokens, latents, indices, final decoder tokens, rec);
the backward re-runs a block's forward from its checkpoint (qkv, attention + log-sum-exp, x_mid, x12) right before
differentiating it.  16 checkpoints x 268 MB at B = 256 — 4.3 GB of the 180 GB.  When memory allows (`keep_attention`,
decided per call from the free HBM) the forward writes every block's qkv, attention output (bf16 + fp32), log-sum-exp,
x_mid and output to fresh tensors — 2.2 GB per block at B = 256; the block input is then its own checkpoint (no copy)
and the backward recomputes only x12 (the attention recomputation was the most expensive part: 1.26 of 2.0 ms per block).

Backward of one block (dgrad = pm_gemm_bf16 against the transposed weight, wgrad = pm_wgrad_bf16):
    dh   = dx W3                     dW3 = dx^T h        db3 = colsum(dx)
    d12  = swiglu'(x12, dh)          dW12 = d12^T n2     db12 = colsum(d12)        n2 = LN2(x_mid)
    dxm  = dx + LN2'(d12 W12)        dWo = dxm^T ao      dbo = colsum(dxm)
    dqkv = attention'(dxm Wo)        dWqkv = dqkv^T n1                             n1 = LN1(x_in)
    dx   = dxm + LN1'(dqkv Wqkv)
"""
from __future__ import annotations

import torch

from . import ops
from .engine import LN_EPS, LOG2E, _RowStats, _fingerprint, run_blocks
from .ops import PM_OUT_F32, PM_OUT_UNPATCH

LN2 = 0.6931471805599453


def _t_bf16(w):
    return w.detach().to(torch.bfloat16).t().contiguous()


class _BlockBwd:
    """Transposed / unfolded operands of one block for the backward pass, and the maps from packed gradients back to
    the reference parameter layout."""

    def __init__(self, layer, fwd_blk):
        a1, ff = layer.attn1, layer.ffnet
        dev = a1.to_q.weight.device
        # the forward runs on PRE-SCALED queries (engine._Block.w_qkv_ps: scale * log2(e) folded into the to_q rows), so the
        # backward differentiates with respect to q' = c q: dn = dqkv' W'^T uses the same scaled rows, and the to_q rows of the
        # weight gradient are multiplied by c afterwards (qscale)
        self.qscale = float(a1.scale) * LOG2E
        wqkv = torch.cat([a1.to_q.weight.detach().float() * self.qscale, a1.to_k.weight.detach().float(), a1.to_v.weight.detach().float()], dim=0)
        self.w_qkv_t = _t_bf16(wqkv)                                   # [D, 3 inner]
        self.w_o_t = _t_bf16(a1.to_out[0].weight)                     # [inner, D]
        h = ff.w12.weight.shape[0] // 2
        hp = fwd_blk.hp
        self.h, self.hp = h, hp
        # packed row r of the tile-interleaved w12 <- reference row src[r] (or -1 for padding)
        T = hp // 128
        j = torch.arange(128, device=dev)
        gate = (torch.arange(T, device=dev)[:, None] * 128 + j[None, :])          # [T, 128] hidden index
        src = torch.stack([gate, gate + h], dim=1).reshape(-1)                    # value rows live at h + hidden index
        valid = torch.stack([gate < h, gate < h], dim=1).reshape(-1)
        self.packed_rows = torch.nonzero(valid).reshape(-1)                        # packed rows that are real
        self.orig_rows = src[valid]                                                # their reference row
        w12p = torch.zeros(2 * hp, ff.w12.weight.shape[1], device=dev, dtype=torch.bfloat16)
        w12p[self.packed_rows] = ff.w12.weight.detach().to(torch.bfloat16)[self.orig_rows]
        self.w_12_t = w12p.t().contiguous()                            # [D, 2 hp]
        w3p = torch.zeros(ff.w3.weight.shape[0], hp, device=dev, dtype=torch.bfloat16)
        w3p[:, :h] = ff.w3.weight.detach().to(torch.bfloat16)
        self.w_3_t = w3p.t().contiguous()                              # [hp, D]
        self.g1 = layer.norm1.weight.detach().float().contiguous()
        self.b1 = layer.norm1.bias.detach().float().contiguous()
        self.g2 = layer.norm2.weight.detach().float().contiguous()
        self.b2 = layer.norm2.bias.detach().float().contiguous()


class Stage1TrainEngine:
    """Forward-with-checkpoints and backward of VQModel.forward; shares the inference engine's packed operands."""

    def __init__(self, model):
        self.model = model
        self.eng = model.engine()
        self._fp = None
        self.ws = self.eng.ws            # one activation workspace for forward and backward
        self.keep_attention = None       # None = decide from free memory; True / False force the policy

    def _ensure_packed(self):
        self.eng._ensure_packed()
        fp = _fingerprint(self.model)
        if fp == self._fp:
            return
        m, e = self.model, self.eng
        with torch.no_grad():
            self.enc_bwd = [_BlockBwd(l, b) for l, b in zip(m.encoder.transformer.layers, e.enc_blocks)]
            self.dec_bwd = [_BlockBwd(l, b) for l, b in zip(m.decoder.transformer.layers, e.dec_blocks)]
            wp = m.prev_quant.weight.detach().float()                                  # [32, D]
            self.w_prev_t2 = torch.cat([wp, wp], dim=0).to(torch.bfloat16).t().contiguous()   # [D, 64] against [hi | lo] of dz
            self.w_post_t = _t_bf16(m.post_quant.weight)                                 # [32, D]
            dec = m.decoder
            P, Cc = dec.patch_size, dec.out_channels
            self.proj_perm = torch.arange(P * P * Cc, device=wp.device).view(P, P, Cc).permute(2, 0, 1).reshape(-1)
            self.w_proj_chw_t = dec.proj.weight.detach()[self.proj_perm].to(torch.bfloat16).t().contiguous()   # [D, 192]
            self.gf = dec.norm.weight.detach().float().contiguous()
            self.bf = dec.norm.bias.detach().float().contiguous()
        self._fp = fp

    # ---------------------------------------------------------------------------------------- forward
    def forward(self, img):
        """img fp32 NCHW -> (rec fp32 NCHW, loss fp32 [], saved dict).  Same kernels and order as Stage1Engine.encode
        + decode (vqmodel.py:32-36), plus the per-block checkpoints."""
        self._ensure_packed()
        m, e = self.model, self.eng
        enc, dec = m.encoder, m.decoder
        if not img.is_cuda:
            raise RuntimeError("paintmind_b200: input must be a CUDA tensor (no CPU fallback)")
        img = img.detach().float().contiguous()
        B, C, H, W = img.shape
        if H != enc.image_size or W != enc.image_size or C != enc.in_channels:
            raise RuntimeError(f"expected input [B,{enc.in_channels},{enc.image_size},{enc.image_size}], got {tuple(img.shape)}")
        g = H // 8
        N, D = g * g, enc.dim
        M = B * N
        dev = img.device
        ws = e.ws
        sv = {"B": B, "N": N}
        patches = torch.empty(M, C * 64, device=dev, dtype=torch.bfloat16)
        x0 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
        x = ws.get("x", (M, D), torch.bfloat16, dev)
        st = _RowStats(ws, M, D, dev)
        ops.patchify8(img, patches)
        ops.gemm(patches, e.w_pe, x0, **e.enc_pos)
        ops.layernorm(x0, gamma=e.pre_g, beta=e.pre_b, y=x, stats=st.finished())
        sv["patches"], sv["x0"] = patches, x0
        keep = self._keep_policy(M, enc.dim, len(e.enc_blocks) + len(e.dec_blocks), dev)
        sv["enc_ckpt"], sv["enc_attn"] = [], []
        x = x.clone() if keep else x                 # keep=True: blocks write fresh tensors, their inputs ARE the checkpoints
        for blk in e.enc_blocks:
            sv["enc_ckpt"].append(x if keep else x.clone())
            x, kept = self._block_forward(blk, x, st, B, N, keep)
            sv["enc_attn"].append(kept)
        sv["x_enc"] = x if keep else x.clone()
        z = torch.empty(M, m.quantize.e_dim, device=dev, dtype=torch.float32)
        ops.gemm(x, e.w_prev, z, bias=e.b_prev, out_mode=PM_OUT_F32, bn=32)
        r = m.quantize.quantize_2d(z, want_split=True)
        sv["z"], sv["idx"], sv["zs"] = z, r["idx"], r["zq_split"]
        self.last_indices = r["idx"].view(B, N)      # code indices of the latest training forward (usage monitoring, tests)
        loss = (r["sse"] * ((1.0 + m.quantize.beta) / (M * m.quantize.e_dim))).to(torch.float32).reshape(())
        # decoder
        Dd = dec.dim
        xd = ws.get("x", (M, Dd), torch.bfloat16, dev)
        st = _RowStats(ws, M, Dd, dev)
        ops.gemm(r["zq_split"], e.w_post, xd, bias=e.b_post, stats_out=st.produce(), **e.dec_pos)
        sv["dec_ckpt"], sv["dec_attn"] = [], []
        xd = xd.clone() if keep else xd
        for blk in e.dec_blocks:
            sv["dec_ckpt"].append(xd if keep else xd.clone())
            xd, kept = self._block_forward(blk, xd, st, B, N, keep)
            sv["dec_attn"].append(kept)
        sv["x_dec"] = xd if keep else xd.clone()
        rec = torch.empty(B, dec.out_channels, dec.image_size, dec.image_size, device=dev, dtype=torch.float32)
        ops.gemm(xd, e.w_proj_chw, rec, bias=e.b_proj_chw, colsum=e.cs_proj_chw, out_mode=PM_OUT_UNPATCH, patch=8, channels=3,
                 grid=dec.image_size // 8, **st.consume())
        sv["rec"] = rec
        return rec, loss, sv

    def _keep_policy(self, M, D, n_blocks, dev):
        if self.keep_attention is not None:
            return bool(self.keep_attention)
        per_block = M * D * (3 * 2 + 2 + 4 + 2 + 2) + M * 8 * 4    # qkv + ao (bf16) + ao (fp32) + x_mid + block output, lse; inner == D
        free, _ = torch.cuda.mem_get_info(dev)
        return n_blocks * per_block < 0.4 * free

    def _block_forward(self, blk, x, st, B, N, keep):
        """One pre-LN block (same kernels as engine.run_blocks) -> (block output, kept tensors).  keep=False: in place on x,
        nothing kept.  keep=True: qkv / attention output / log-sum-exp / x_mid / the block output go to fresh tensors: x itself
        is left intact (it is the block's checkpoint — no copy), and the backward pass recomputes only x12."""
        if not keep:
            run_blocks([blk], x, st, B, N, self.eng.ws)
            return x, None
        M, D = x.shape
        dev = x.device
        inner, H = blk.inner, blk.heads
        qkv = torch.empty(M, 3 * inner, device=dev, dtype=torch.bfloat16)
        ao = torch.empty(M, inner, device=dev, dtype=torch.bfloat16)
        ao32 = torch.empty(B, N, inner, device=dev, dtype=torch.float32)
        lse = ops.lse_buffer(B, H, N, dev)
        ops.gemm(x, blk.w_qkv_ps, qkv, bias=blk.b_qkv_ps, colsum=blk.cs_qkv_ps, **st.consume())
        q3 = qkv.view(B, N, 3 * inner)
        ops.attention_train(q3[..., :inner], q3[..., inner:2 * inner], q3[..., 2 * inner:], ao.view(B, N, inner), H, blk.scale, lse, ao32,
                            prescaled=True)
        x_mid = torch.empty_like(x)
        ops.gemm(ao, blk.w_o, x_mid, bias=blk.b_o, res=x, stats_out=st.produce())
        h = self.eng.ws.get("h", (M, blk.hp), torch.bfloat16, dev)
        ops.gemm(x_mid, blk.w_12, h, bias=blk.b_12, colsum=blk.cs_12, swiglu=True, **st.consume())
        x_out = torch.empty_like(x)
        ops.gemm(h, blk.w_3, x_out, bias=blk.b_3, res=x_mid, stats_out=st.produce())
        return x_out, (qkv, ao, ao32, lse, x_mid)

    # ---------------------------------------------------------------------------------------- backward
    def _block_backward(self, blk, bw, x_in, dx, B, N, grads, prefix, kept=None):
        """dx (bf16 [M, D], gradient of the block OUTPUT) is replaced by the gradient of the block INPUT; parameter
        gradients are written into `grads` under the reference names."""
        ws = self.ws
        M, D = x_in.shape
        dev = x_in.device
        inner, hp, H = blk.inner, blk.hp, blk.heads
        bf = torch.bfloat16
        # ---- recompute the block's forward from its checkpoint ----
        stats = ws.get("stats", (M, 2), torch.float32, dev)
        x12 = ws.get("x12", (M, 2 * hp), bf, dev)
        if kept is not None:
            qkv, ao, ao32, lse, x_mid = kept
            q3 = qkv.view(B, N, 3 * inner)
        else:
            x_mid = ws.get("x_mid", (M, D), bf, dev)
            qkv = ws.get("qkv", (M, 3 * inner), bf, dev)
            ao = ws.get("ao", (M, inner), bf, dev)
            lse = ws.get("lse", (B, H, (N + 127) // 128 * 128), torch.float32, dev)[:, :, :N]
            ao32 = ws.get("ao32", (B, N, inner), torch.float32, dev)
            ops.layernorm(x_in, stats=stats)
            ops.gemm(x_in, blk.w_qkv_ps, qkv, bias=blk.b_qkv_ps, colsum=blk.cs_qkv_ps, stats=stats)
            q3 = qkv.view(B, N, 3 * inner)
            ops.attention_train(q3[..., :inner], q3[..., inner:2 * inner], q3[..., 2 * inner:], ao.view(B, N, inner), H, blk.scale, lse, ao32,
                                prescaled=True)
            ops.gemm(ao, blk.w_o, x_mid, bias=blk.b_o, res=x_in)
        ops.layernorm(x_mid, stats=stats)
        ops.gemm(x_mid, blk.w_12, x12, bias=blk.b_12, colsum=blk.cs_12, stats=stats)          # plain store: gate | value tiles
        # ---- feed-forward branch ----
        dh = ws.get("dh", (M, hp), bf, dev)
        h = ws.get("h", (M, hp), bf, dev)
        d12 = ws.get("d12", (M, 2 * hp), bf, dev)
        nbuf = ws.get("n", (M, D), bf, dev)
        dn = ws.get("dn", (M, D), bf, dev)
        dx_mid = ws.get("dx_mid", (M, D), bf, dev)
        f32 = dict(device=dev, dtype=torch.float32)
        b3 = torch.empty(D, **f32)
        ops.colsum(dx, b3)
        ops.gemm(dx, bw.w_3_t, dh, bn=256 if M >= 256 else 0)      # 1408 = 5.5 x 256: CTA pairs with a ragged last tile beat bn = 128
        b12 = torch.empty(2 * hp, **f32)
        ops.swiglu_bwd(x12, dh, h, d12, b12)                 # + bias gradient of w12 (column sums of d12) in the same pass
        w3 = torch.empty(D, hp, **f32)
        ops.wgrad(dx, h, w3)
        ops.layernorm(x_mid, gamma=bw.g2, beta=bw.b2, y=nbuf)
        w12 = torch.empty(2 * hp, D, **f32)
        ops.wgrad(d12, nbuf, w12)
        ops.gemm(d12, bw.w_12_t, dn)
        gb2 = torch.empty(2, D, **f32)
        ops.layernorm_bwd(dn, x_mid, bw.g2, dx_mid, gb2, dres=dx, eps=LN_EPS)
        # ---- attention branch ----
        bo = torch.empty(D, **f32)
        ops.colsum(dx_mid, bo)
        wo = torch.empty(D, inner, **f32)
        ops.wgrad(dx_mid, ao, wo)
        dao = ws.get("dao", (M, inner), bf, dev)
        ops.gemm(dx_mid, bw.w_o_t, dao)
        dqkv = ws.get("dqkv", (M, 3 * inner), bf, dev)
        d3 = dqkv.view(B, N, 3 * inner)
        ops.attention_bwd(q3[..., :inner], q3[..., inner:2 * inner], q3[..., 2 * inner:], ao32, dao.view(B, N, inner), lse,
                          d3[..., :inner], d3[..., inner:2 * inner], d3[..., 2 * inner:], H, LN2)
        # (q is pre-scaled: the natural-log logits are ln2 * q' k^T, so `scale` = ln 2 makes the kernel's exp2 argument q' k^T - lse
        #  and its dq the gradient with respect to q'; dk and dv are the true gradients)
        ops.layernorm(x_in, gamma=bw.g1, beta=bw.b1, y=nbuf)
        wqkv = torch.empty(3 * inner, D, **f32)
        ops.wgrad(dqkv, nbuf, wqkv)
        wqkv[:inner].mul_(bw.qscale)                         # d/d to_q.weight = c * d/d (c to_q.weight)
        ops.gemm(dqkv, bw.w_qkv_t, dn)
        gb1 = torch.empty(2, D, **f32)
        ops.layernorm_bwd(dn, x_in, bw.g1, dx, gb1, dres=dx_mid, eps=LN_EPS)
        # ---- back to the reference parameter layout ----
        hdim = bw.h
        g12 = torch.empty(2 * hdim, D, **f32)
        g12[bw.orig_rows] = w12[bw.packed_rows]
        gb12 = torch.empty(2 * hdim, **f32)
        gb12[bw.orig_rows] = b12[bw.packed_rows]
        grads[prefix + "norm1.weight"], grads[prefix + "norm1.bias"] = gb1[0], gb1[1]
        grads[prefix + "attn1.to_q.weight"] = wqkv[:inner]
        grads[prefix + "attn1.to_k.weight"] = wqkv[inner:2 * inner]
        grads[prefix + "attn1.to_v.weight"] = wqkv[2 * inner:]
        grads[prefix + "attn1.to_out.0.weight"], grads[prefix + "attn1.to_out.0.bias"] = wo, bo
        grads[prefix + "norm2.weight"], grads[prefix + "norm2.bias"] = gb2[0], gb2[1]
        grads[prefix + "ffnet.w12.weight"], grads[prefix + "ffnet.w12.bias"] = g12, gb12
        grads[prefix + "ffnet.w3.weight"], grads[prefix + "ffnet.w3.bias"] = w3[:, :hdim], b3

    def backward(self, sv, d_rec, d_loss):
        """Gradients of every VQModel parameter (dict keyed like state_dict) from d rec [B,3,H,W] and d loss []."""
        m, e = self.model, self.eng
        enc, dec = m.encoder, m.decoder
        B, N = sv["B"], sv["N"]
        M = B * N
        dev = sv["rec"].device
        ws = self.ws
        bf = torch.bfloat16
        f32 = dict(device=dev, dtype=torch.float32)
        grads = {}
        D = dec.dim
        # ---- proj + un-patchify + clamp (vqmodel.py:30, layers.py:148-150) ----
        x_dec = sv["x_dec"]
        dx = ws.get("dx", (M, D), bf, dev)
        nbuf = ws.get("n", (M, D), bf, dev)
        dn = ws.get("dn", (M, D), bf, dev)
        n_out = dec.out_channels * 64
        if d_rec is not None:
            dy = ws.get("dy_proj", (M, n_out), bf, dev)
            ops.unpatchify8_bwd(d_rec.detach().float().contiguous(), sv["rec"], dy)
            ops.layernorm(x_dec, gamma=self.gf, beta=self.bf, y=nbuf)
            wproj = torch.empty(n_out, D, **f32)
            bproj = torch.empty(n_out, **f32)
            ops.wgrad(dy, nbuf, wproj)
            ops.colsum(dy, bproj)
            gw = torch.empty_like(wproj)
            gw[self.proj_perm] = wproj
            gb = torch.empty_like(bproj)
            gb[self.proj_perm] = bproj
            grads["decoder.proj.weight"], grads["decoder.proj.bias"] = gw, gb
            ops.gemm(dy, self.w_proj_chw_t, dn)
            gbf = torch.empty(2, D, **f32)
            ops.layernorm_bwd(dn, x_dec, self.gf, dx, gbf, eps=LN_EPS)
            grads["decoder.norm.weight"], grads["decoder.norm.bias"] = gbf[0], gbf[1]
            # ---- decoder blocks, last to first ----
            for li in reversed(range(len(e.dec_blocks))):
                self._block_backward(e.dec_blocks[li], self.dec_bwd[li], sv["dec_ckpt"][li], dx, B, N, grads,
                                     f"decoder.transformer.layers.{li}.", sv["dec_attn"][li])
                sv["dec_ckpt"][li] = sv["dec_attn"][li] = None          # release as the backward pass retreats
            # ---- post_quant + decoder position embedding (vqmodel.py:28, layers.py:146) ----
            gpos = torch.empty(N * D, **f32)
            ops.colsum(dx.view(B, N * D), gpos)
            grads["decoder.position_embedding"] = gpos.view(1, N, D)
            bpost = torch.empty(D, **f32)
            ops.colsum(dx, bpost)
            wpost = torch.empty(D, 64, **f32)
            ops.wgrad(dx, sv["zs"], wpost)
            grads["post_quant.weight"], grads["post_quant.bias"] = wpost[:, :32] + wpost[:, 32:], bpost
            d_zq = ws.get("d_zq", (M, 32), torch.float32, dev)
            ops.gemm(dx, self.w_post_t, d_zq, out_mode=PM_OUT_F32, bn=32)
        else:
            d_zq = None
        # ---- quantizer: straight-through + both loss terms (quantize.py:19,29-36) ----
        E = m.quantize.embedding.weight.detach().float().contiguous()
        dzs = ws.get("dz_split", (M, 64), bf, dev)
        dl = d_loss.detach().float().reshape(1).contiguous() if d_loss is not None else None
        dE = torch.zeros_like(E) if dl is not None else None       # the codebook only hears from the loss terms (quantize.py:33,36)
        ops.vq_bwd(sv["z"], sv["idx"], E, d_zq, dl, m.quantize.beta, dz=None, dz_split=dzs, dE=dE)
        if dE is not None:
            grads["quantize.embedding.weight"] = dE
        # ---- prev_quant (vqmodel.py:23) ----
        De = enc.dim
        wprev = torch.empty(64, De, **f32)
        ops.wgrad(dzs, sv["x_enc"], wprev)
        bprev = torch.empty(64, **f32)
        ops.colsum(dzs, bprev)
        grads["prev_quant.weight"], grads["prev_quant.bias"] = wprev[:32] + wprev[32:], bprev[:32] + bprev[32:]
        dxe = ws.get("dx", (M, De), bf, dev)
        ops.gemm(dzs, self.w_prev_t2, dxe)
        # ---- encoder blocks ----
        for li in reversed(range(len(e.enc_blocks))):
            self._block_backward(e.enc_blocks[li], self.enc_bwd[li], sv["enc_ckpt"][li], dxe, B, N, grads,
                                 f"encoder.transformer.layers.{li}.", sv["enc_attn"][li])
            sv["enc_ckpt"][li] = sv["enc_attn"][li] = None
        # ---- norm_pre, position embedding, patch embedding (layers.py:107-109) ----
        dx0 = ws.get("dn", (M, De), bf, dev)
        gbp = torch.empty(2, De, **f32)
        ops.layernorm_bwd(dxe, sv["x0"], e.pre_g, dx0, gbp, eps=LN_EPS)
        grads["encoder.norm_pre.weight"], grads["encoder.norm_pre.bias"] = gbp[0], gbp[1]
        gpos = torch.empty(N * De, **f32)
        ops.colsum(dx0.view(B, N * De), gpos)
        grads["encoder.position_embedding"] = gpos.view(1, N, De)
        kpe = sv["patches"].shape[1]
        wpe = torch.empty(De, kpe, **f32)
        ops.wgrad(dx0, sv["patches"], wpe)
        grads["encoder.to_patch_embedding.0.weight"] = wpe.view(De, enc.in_channels, 8, 8)
        return grads


class _VQGANStep(torch.autograd.Function):
    """(rec, loss) = VQModel.forward(img) with gradients for every parameter (not for img: no caller needs it —
    the discriminator's gradient penalty differentiates the discriminator, utils/trainer.py:153-169)."""

    @staticmethod
    def forward(ctx, model, img, names, *params):
        te = model.train_engine()
        rec, loss, sv = te.forward(img)
        ctx.model, ctx.sv, ctx.names = model, sv, names
        ctx.set_materialize_grads(False)       # an unused output arrives as None, not as a tensor of zeros
        return rec, loss

    @staticmethod
    def backward(ctx, d_rec, d_loss):
        te = ctx.model.train_engine()
        with torch.no_grad():
            grads = te.backward(ctx.sv, d_rec, d_loss)
        ctx.sv = None
        out = []
        for i, n in enumerate(ctx.names):
            g = grads.get(n) if ctx.needs_input_grad[3 + i] else None
            out.append(g)
        return (None, None, None, *out)


def vqgan_forward_with_grad(model, img):
    if img.requires_grad:
        raise NotImplementedError("paintmind_b200: VQModel.forward produces parameter gradients only; a gradient w.r.t. the input "
                                  "image is not computed (no caller of the reference needs one) — detach the image")
    names, params = zip(*[(n, p) for n, p in model.named_parameters()])
    return _VQGANStep.apply(model, img, names, *params)
