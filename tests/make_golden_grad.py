"""Golden fixture for the generator BACKWARD path from the unmodified reference under torch
autograd.  BUILD-CONTAINER ONLY (imports /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/make_golden_grad.py

The reference differentiates `rec, codebook_loss = vqvae(img)` (utils/trainer.py:205-217).  LPIPS and the discriminator
are off the path, so the objective here is the part of the generator loss that only involves the path:
    L = codebook_loss + F.l1_loss(rec, img) + F.mse_loss(rec, img)          (trainer.py:207-208,215)
evaluated in fp32 on CPU on the seeded synthetic weights / images.  The 222 gradient tensors hold 25 M numbers, so the
fixture stores, per parameter, the gradient's L2 norm and its values at 512 seeded positions (grad_sample_positions);
small tensors (<= 4096 elements) are stored whole.  Also recorded: the same gradients from the reference under
`torch.autocast('cpu', bfloat16)` as relative-L2 distances to the fp32 ones — the reference's own mixed-precision
noise, which is what the bf16 CUDA path is held against."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")

from oracle.ref_loader import load_reference  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402
from grad_sampling import SMALL, grad_sample_positions  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def objective(rec, closs, img):
    return closs + F.l1_loss(rec, img) + F.mse_loss(rec, img)


def run(model, img, autocast=False):
    model.zero_grad(set_to_none=True)
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            rec, closs = model(img)
            L = objective(rec.float(), closs, img)
    else:
        rec, closs = model(img)
        L = objective(rec, closs, img)
    L.backward()
    return rec.detach(), closs.detach(), L.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    pm = load_reference()
    from paintmind.stage1 import VQModel
    for cfg_name, batch, seed, out_name in (("vit-s-vqgan", 2, 0, "stage1_grad_vit_s.npz"), ("vit-tiny-test", 3, 7, "stage1_grad_tiny.npz")):
        cfg = ver2cfg[cfg_name]
        sd = synthetic.make_vqgan_state_dict(cfg, seed=seed)
        model = VQModel(pm.Config(cfg)).train()
        res = model.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        img = synthetic.make_images(batch, cfg["enc"]["image_size"], seed=seed + 200)
        rec, closs, L, g32 = run(model, img)
        with torch.no_grad():
            _, _, idx = model.encode(img)
        out = {"loss": np.float64(L), "codebook_loss": np.float64(closs), "rec_abs_sum": np.float64(rec.double().abs().sum()),
               "indices": idx.numpy().astype(np.int32), "batch": np.int64(batch), "seed": np.int64(seed)}
        names = sorted(g32.keys())
        out["names"] = np.array(names)
        out["norms"] = np.array([float(g32[n].double().norm()) for n in names])
        for i, n in enumerate(names):
            g = g32[n].reshape(-1)
            if g.numel() <= SMALL:
                out[f"g{i}"] = g.numpy().astype(np.float32)
            else:
                out[f"g{i}"] = g[torch.from_numpy(grad_sample_positions(n, g.numel()))].numpy().astype(np.float32)
        _, _, Lb, gbf = run(model, img, autocast=True)
        out["autocast_rel_l2"] = np.array([float((gbf[n].float() - g32[n]).norm() / g32[n].norm().clamp_min(1e-30)) for n in names])
        out["autocast_loss"] = np.float64(Lb)
        keys = sorted(sd.keys())
        picks = keys[:: max(1, len(keys) // 8)]
        out["weight_keys"] = np.array(picks)
        out["weight_sums"] = np.array([float(sd[k].double().abs().sum()) for k in picks])
        out["img_abs_sum"] = np.float64(img.double().abs().sum())
        np.savez_compressed(GOLD / out_name, **out)
        ar = out["autocast_rel_l2"]
        print(f"{out_name}: loss {float(L):.6f} (autocast {float(Lb):.6f}); reference bf16-autocast vs fp32 gradient rel-L2: "
              f"median {np.median(ar):.4f} max {ar.max():.4f} ({names[int(ar.argmax())]}); {os.path.getsize(GOLD / out_name) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
