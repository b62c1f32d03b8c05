"""Generate golden fixtures from the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/make_golden.py [stage1|vq|stage2|all]

Imports /root/reference read-only through oracle/ref_loader.py, loads the seeded synthetic
state_dict (paintmind_b200/utils/synthetic.py) into the reference modules, runs the reference's own
public API on CPU fp32 and stores its outputs as small .npz files under tests/golden/.  The GPU box
has no /root/reference: tests there regenerate the same seeded weights/inputs and compare against
these files.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")

from oracle.ref_loader import load_reference  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

GOLD = ROOT / "tests" / "golden"
GOLD.mkdir(parents=True, exist_ok=True)


def weight_checksums(sd):
    keys = sorted(sd.keys())
    picks = keys[:: max(1, len(keys) // 8)]
    return {k: float(sd[k].double().abs().sum()) for k in picks}


def ref_vqgan(pm, cfg_name, sd):
    from paintmind.stage1 import VQModel
    cfg = pm.Config(ver2cfg[cfg_name])
    model = VQModel(cfg).eval()
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model


def stage1_fixture(pm, cfg_name, batch, seed, out_name, rec_stride):
    cfg = ver2cfg[cfg_name]
    sd = synthetic.make_vqgan_state_dict(cfg, seed=seed)
    model = ref_vqgan(pm, cfg_name, sd)
    x = synthetic.make_images(batch, cfg["enc"]["image_size"], seed=seed + 100)
    with torch.no_grad():
        tokens = model.encoder(x)
        z_pre = model.prev_quant(tokens)
        z_q, loss, idx = model.encode(x)
        # distances for the top-2 gap, with the reference's own formula (quantize.py:19-26)
        zn = torch.nn.functional.normalize(z_pre, dim=-1).view(-1, cfg["embed_dim"])
        en = torch.nn.functional.normalize(model.quantize.embedding.weight, dim=-1)
        d = (zn ** 2).sum(1, keepdim=True) + (en ** 2).sum(1) - 2 * torch.einsum("bd,nd->bn", zn, en)
        top2 = d.topk(2, dim=1, largest=False).values
        gap = (top2[:, 1] - top2[:, 0]).view(idx.shape)
        rec = model.decode(z_q)
        rec_from_idx = model.decode_from_indice(idx)
        # un-clamped decoder output for a stricter comparison than the saturating clamp allows
        pre = model.decoder(model.post_quant(z_q))
    hist = torch.bincount(idx.view(-1), minlength=cfg["n_embed"])
    np.savez_compressed(
        GOLD / out_name,
        cfg_name=cfg_name, batch=batch, seed=seed,
        weight_keys=np.array(list(weight_checksums(sd).keys())),
        weight_sums=np.array(list(weight_checksums(sd).values())),
        x_sum=float(x.double().sum()),
        tokens_sub=tokens[:, ::8, ::8].numpy().astype(np.float32),
        z_pre=z_pre.numpy().astype(np.float32),
        z_q=z_q.numpy().astype(np.float32),
        loss=float(loss),
        idx=idx.numpy().astype(np.int16),
        gap=gap.numpy().astype(np.float32),
        hist_nonzero_bins=hist.nonzero().view(-1).numpy().astype(np.int16),
        hist_nonzero_counts=hist[hist > 0].numpy().astype(np.int32),
        rec_sub=rec[:, :, ::rec_stride, ::rec_stride].numpy().astype(np.float32),
        pre_sub=pre[:, :, ::rec_stride, ::rec_stride].numpy().astype(np.float32),
        rec_mean=float(rec.double().mean()), rec_absmean=float(rec.double().abs().mean()),
        rec_sat_frac=float((rec.abs() >= 1.0).double().mean()),
        rec_idx_maxdiff=float((rec - rec_from_idx).abs().max()),
        rec_stride=rec_stride,
    )
    print(f"{out_name}: loss={float(loss):.6f} used_codes={(hist > 0).sum().item()} min_gap={gap.min().item():.3g} "
          f"sat={float((rec.abs() >= 1.0).double().mean()):.3f} rec_vs_idx={float((rec - rec_from_idx).abs().max()):.3g}")


def vq_fixture(pm):
    """BASELINE config 2: 65,536 l2-normalised latents vs the 8192 x 32 codebook."""
    from paintmind.stage1.quantize import VectorQuantizer
    g = torch.Generator().manual_seed(0)
    z = torch.nn.functional.normalize(torch.randn(65536, 32, generator=g), dim=-1)
    E = torch.randn(8192, 32, generator=g)
    vq = VectorQuantizer(8192, 32, 0.25)
    vq.embedding.weight.data.copy_(E)
    with torch.no_grad():
        z_q, loss, idx = vq(z.view(64, 1024, 32))
        en = torch.nn.functional.normalize(E, dim=-1)
        d = (z ** 2).sum(1, keepdim=True) + (en ** 2).sum(1) - 2 * torch.einsum("bd,nd->bn", z, en)
        top2 = d.topk(2, dim=1, largest=False).values
        gap = top2[:, 1] - top2[:, 0]
        dec = vq.decode_from_indice(idx[:1, :16])
    np.savez_compressed(
        GOLD / "vq_microbench.npz",
        z_sum=float(z.double().sum()), E_sum=float(E.double().sum()),
        idx=idx.view(-1).numpy().astype(np.int16), loss=float(loss),
        gap=gap.numpy().astype(np.float32),
        z_q_head=z_q.view(-1, 32)[:64].numpy().astype(np.float32),
        dec_head=dec.view(-1, 32).numpy().astype(np.float32),
    )
    print(f"vq_microbench: loss={float(loss):.6f} min_gap={gap.min().item():.3g} n(gap<1e-5)={(gap < 1e-5).sum().item()}")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.set_num_threads(os.cpu_count() or 1)
    pm = load_reference()
    if what in ("stage1", "all"):
        stage1_fixture(pm, "vit-tiny-test", batch=2, seed=7, out_name="stage1_tiny.npz", rec_stride=1)
        stage1_fixture(pm, "vit-s-vqgan", batch=2, seed=0, out_name="stage1_vit_s.npz", rec_stride=4)
    if what in ("vq", "all"):
        vq_fixture(pm)
    if what in ("stage2", "all"):
        from make_golden_stage2 import stage2_fixtures
        stage2_fixtures(pm)


if __name__ == "__main__":
    main()
