"""Golden fixture at the HEADLINE configuration (BASELINE configs[2]: vit-s-vqgan, batch 256) from the UNMODIFIED
reference — build container only.

    PYTHONDONTWRITEBYTECODE=1 python tests/make_golden_headline.py

The reference's VQModel (stage1/vqmodel.py:21-30) is run on CPU fp32 over the 256 seeded images in chunks of 16 (the path
has no cross-sample operation, so chunking does not change any value; one batch-256 call would need ~17 GB for the
attention and distance matrices) and the outputs are reduced to what the GPU test needs (tests/test_gpu_headline.py):
    idx [256, 1024] int16, gap [256, 1024] fp16 (rounded UP: never under-states a gap), the reference latents of every
    16th token (fp32, for the per-token Lipschitz rule), loss, the usage histogram, and of the reconstruction decoded from
    the reference's own latents: per-image mean / abs-mean (fp64) and a [::8, ::8] pixel sample (fp16).
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")

from oracle.ref_loader import load_reference  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

BATCH, CHUNK, SEED, IMG_SEED, TOK_STRIDE, PIX_STRIDE = 256, 16, 0, 1100, 16, 8


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    pm = load_reference()
    from paintmind.stage1 import VQModel
    cfg = ver2cfg["vit-s-vqgan"]
    sd = synthetic.make_vqgan_state_dict(cfg, seed=SEED)
    model = VQModel(pm.Config(cfg)).eval()
    model.load_state_dict(sd, strict=True)
    x = synthetic.make_images(BATCH, 256, seed=IMG_SEED)
    en = torch.nn.functional.normalize(model.quantize.embedding.weight.detach(), dim=-1)
    idx_all, gap_all, zsub, rec_sub, rec_mean, rec_abs, sse = [], [], [], [], [], [], 0.0
    t0 = time.time()
    with torch.no_grad():
        for c in range(0, BATCH, CHUNK):
            xc = x[c:c + CHUNK]
            z_pre = model.prev_quant(model.encoder(xc))
            z_q, loss, idx = model.encode(xc)
            zn = torch.nn.functional.normalize(z_pre, dim=-1).view(-1, cfg["embed_dim"])
            d = (zn ** 2).sum(1, keepdim=True) + (en ** 2).sum(1) - 2 * torch.einsum("bd,nd->bn", zn, en)   # quantize.py:22-26
            top2 = d.topk(2, dim=1, largest=False).values
            gap_all.append((top2[:, 1] - top2[:, 0]).view(idx.shape))
            idx_all.append(idx)
            zsub.append(z_pre[:, ::TOK_STRIDE].contiguous())
            sse += float(((z_q - zn.view_as(z_q)) ** 2).double().sum())
            rec = model.decode(z_q)
            rec_sub.append(rec[:, :, ::PIX_STRIDE, ::PIX_STRIDE].contiguous())
            rec_mean.append(rec.double().mean(dim=(1, 2, 3)))
            rec_abs.append(rec.double().abs().mean(dim=(1, 2, 3)))
            print(f"  {c + CHUNK}/{BATCH} images, {time.time() - t0:.0f} s", flush=True)
    idx = torch.cat(idx_all)
    gap = torch.cat(gap_all)
    gap16 = gap.to(torch.float16)
    gap16 = torch.where(gap16.float() < gap, torch.nextafter(gap16, torch.full_like(gap16, 65504.0)), gap16)   # round UP
    hist = torch.bincount(idx.view(-1), minlength=cfg["n_embed"])
    loss = (1.0 + cfg["beta"]) * sse / (idx.numel() * cfg["embed_dim"])
    keys = sorted(sd.keys())
    picks = keys[:: max(1, len(keys) // 8)]
    out = ROOT / "tests" / "golden" / "headline_vit_s_b256.npz"
    np.savez_compressed(
        out, cfg_name="vit-s-vqgan", batch=BATCH, seed=SEED, img_seed=IMG_SEED, tok_stride=TOK_STRIDE, pix_stride=PIX_STRIDE,
        weight_keys=np.array(picks), weight_sums=np.array([float(sd[k].double().abs().sum()) for k in picks]),
        x_sum=float(x.double().sum()),
        idx=idx.numpy().astype(np.int16), gap=gap16.numpy(), z_pre_sub=torch.cat(zsub).numpy().astype(np.float32),
        loss=float(loss), hist_nonzero_bins=hist.nonzero().view(-1).numpy().astype(np.int16),
        hist_nonzero_counts=hist[hist > 0].numpy().astype(np.int32),
        rec_sub=torch.cat(rec_sub).numpy().astype(np.float16), rec_mean=torch.cat(rec_mean).numpy(), rec_absmean=torch.cat(rec_abs).numpy(),
    )
    print(f"{out.name}: loss={loss:.6f} used_codes={(hist > 0).sum().item()} min_gap={gap.min().item():.3g} "
          f"size={out.stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
