"""Where the reconstruction error of the bf16 path comes from (CPU emulation, minutes).

The decoder of vit-s-vqgan (stage1/layers.py:145-152 + vqmodel.py:27-30) is evaluated on the reference's own latents
(tests/golden/stage1_vit_s.npz) in fp32 and with the roundings of three execution models inserted in torch:
  autocast : what the reference does under torch.autocast(bf16) (utils/trainer.py:187): bf16 matmul operands and bf16 Linear
             outputs, fp32 LayerNorm / softmax / residual stream;
  ours     : the CUDA path's layout: bf16 operands, fp32 accumulation, LayerNorm folded into the next GEMM
             (raw bf16 x as the A operand), bf16 q|k|v, P, attention output and SwiGLU hidden, and a BF16 RESIDUAL STREAM (one
             rounding of x after every residual update);
  ours+f32x: the same with the residual stream kept in fp32 (what a hi+lo / fp32 residual would buy);
  ours+f16x: the same with the residual stream in FP16 (same bytes as bf16, 11 instead of 8 significant bits; the GEMM A operand is
             then the fp16 x itself).
Prints max / mean abs error of the clamped reconstruction against fp32, on all pixels of both images.
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

cfg = ver2cfg["vit-s-vqgan"]
sd = synthetic.make_vqgan_state_dict(cfg, seed=0)
g = np.load(ROOT / "tests" / "golden" / "stage1_vit_s.npz")
zq = torch.from_numpy(g["z_q"])


def bf(t):
    return t.to(torch.bfloat16).float()


def decoder(z, mode):
    r = (lambda t: t) if mode == "fp32" else bf                         # operand rounding
    out_r = bf if mode == "autocast" else (lambda t: t)                 # autocast: Linear returns bf16
    x_r = bf if mode == "ours" else (lambda t: t.to(torch.float16).float()) if mode == "ours+f16x" else (lambda t: t)   # residual-stream rounding
    dcfg = cfg["dec"]
    H = dcfg["num_head"]

    def lin(x, w, b=None):
        return out_r(F.linear(r(x), r(w), b))

    x = lin(z, sd["post_quant.weight"], sd["post_quant.bias"]) + sd["decoder.position_embedding"]
    x = x_r(x)
    for i in range(dcfg["depth"]):
        p = f"decoder.transformer.layers.{i}."

        def ln_lin(x, norm, w, b=None):
            gam, bet = sd[p + norm + ".weight"], sd[p + norm + ".bias"]
            if mode.startswith("ours"):
                # folded: rstd * (x_bf16 @ (gamma W)_bf16^T - mu * colsum) + (b + W beta); statistics in fp32 of the fp32 sums
                mu = x.mean(-1, keepdim=True)
                rstd = (x.var(-1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
                wf = bf(w * gam[None, :])
                xa = x if mode == "ours+f16x" else bf(x)          # fp16 residual stream: the A operand is x itself (fp16 x bf16 products are exact in fp32)
                y = rstd * (F.linear(xa, wf) - mu * wf.sum(1)[None, None, :]) + (F.linear(bet[None], w)[0] + (b if b is not None else 0))
                return y
            return lin(F.layer_norm(x, (x.shape[-1],), gam, bet, 1e-5), w, b)

        wqkv = torch.cat([sd[p + "attn1.to_q.weight"], sd[p + "attn1.to_k.weight"], sd[p + "attn1.to_v.weight"]], 0)
        qkv = r(ln_lin(x, "norm1", wqkv))
        B, N, _ = qkv.shape
        q, k, v = [t.view(B, N, H, 64).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        s = (q * 0.125) @ k.transpose(-1, -2) if mode != "autocast" else out_r(r(q * 0.125) @ k.transpose(-1, -2))
        pr = torch.softmax(s, dim=-1)
        if mode.startswith("ours"):
            # un-normalised bf16 P against V, fp32 row sum of the fp32 exponentials
            m = s.max(-1, keepdim=True).values
            e = torch.exp(s - m)
            ao = (bf(e) @ v) / e.sum(-1, keepdim=True)
        else:
            ao = out_r(r(pr) @ v)
        ao = r(ao.transpose(1, 2).reshape(B, N, H * 64))
        x = x_r(lin(ao, sd[p + "attn1.to_out.0.weight"], sd[p + "attn1.to_out.0.bias"]) + x)
        x12 = ln_lin(x, "norm2", sd[p + "ffnet.w12.weight"], sd[p + "ffnet.w12.bias"])
        x1, x2 = x12.chunk(2, dim=-1)
        h = r(F.silu(x1) * x2)
        x = x_r(lin(h, sd[p + "ffnet.w3.weight"], sd[p + "ffnet.w3.bias"]) + x)
    if mode.startswith("ours"):
        gam, bet = sd["decoder.norm.weight"], sd["decoder.norm.bias"]
        mu = x.mean(-1, keepdim=True)
        rstd = (x.var(-1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
        w = sd["decoder.proj.weight"]
        wf = bf(w * gam[None, :])
        xa = x if mode == "ours+f16x" else bf(x)
        y = rstd * (F.linear(xa, wf) - mu * wf.sum(1)[None, None, :]) + (F.linear(bet[None], w)[0] + sd["decoder.proj.bias"])
    else:
        y = F.linear(r(F.layer_norm(x, (x.shape[-1],), sd["decoder.norm.weight"], sd["decoder.norm.bias"], 1e-5)), r(sd["decoder.proj.weight"]),
                     sd["decoder.proj.bias"]).float()
    B = y.shape[0]
    return y.view(B, 32, 32, 8, 8, 3).permute(0, 5, 1, 3, 2, 4).reshape(B, 3, 256, 256).clamp(-1, 1)


with torch.no_grad():
    ref = decoder(zq, "fp32")
    s = int(g["rec_stride"])
    print(f"fp32 emulation vs the reference fixture: max {float((ref[:, :, ::s, ::s] - torch.from_numpy(g['rec_sub'])).abs().max()):.2e}")
    for mode in ("autocast", "ours", "ours+f32x", "ours+f16x"):
        err = (decoder(zq, mode) - ref).abs()
        print(f"{mode:10s} max {float(err.max()):.4f}  mean {float(err.mean()):.5f}  p99.9 {float(err.flatten().kthvalue(int(0.999 * err.numel())).values):.4f}")
