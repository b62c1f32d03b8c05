"""Golden fixtures for the stage-2 TRAINING forward from the unmodified reference:
Pipeline.random_masking / loss / forward (generate.py:78-146).  BUILD-CONTAINER ONLY (imports /root/reference).

The reference draws its masking noise with torch.rand(N, L, device=...) inside random_masking; to make the result
reproducible on another RNG the generator temporarily replaces torch.rand with a function returning the seeded
noise of tests/stage2_inputs.py (the reference source is not modified)."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle.ref_loader import load_reference  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402
from stage2_inputs import loss_inputs, masking_inputs, train_forward_inputs  # noqa: E402

GOLD = ROOT / "tests" / "golden"


class inject_rand:
    def __init__(self, noise):
        self.noise = noise

    def __enter__(self):
        self.orig = torch.rand
        torch.rand = lambda *a, **k: self.noise.clone()
        return self

    def __exit__(self, *exc):
        torch.rand = self.orig


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    pm = load_reference()
    cfg1, cfg2 = ver2cfg["vit-s-vqgan"], ver2cfg["paintmindv1"]
    pipe = pm.create_model(arch="pipeline", version="paintmindv1", pretrained=False).eval()
    sd = {("vqgan." + k): v for k, v in synthetic.make_vqgan_state_dict(cfg1, seed=0).items()}
    sd.update(synthetic.make_stage2_state_dict(cfg2, cfg1, seed=1, context_dim=1024))
    res = pipe.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    out = {}
    with torch.no_grad():
        # ---- random_masking at two ratios ----
        x, noise = masking_inputs()
        for tag, ratio in (("75", 0.75), ("30", 0.3), ("tiny", 0.0001)):
            with inject_rand(noise):
                xm, mask = pipe.random_masking(x, ratio)
            out[f"mask_{tag}"] = np.packbits(mask.numpy().astype(np.uint8), axis=1)
            out[f"xm_{tag}_sum"] = np.float64(xm.double().sum())
            out[f"xm_{tag}_head"] = xm[:, :16].numpy().astype(np.float32)
        # ---- loss ----
        logits, label, masks = loss_inputs()
        out["loss_unit"] = np.float32(pipe.loss(logits, label, masks))
        out["loss_unit_allmask"] = np.float32(pipe.loss(logits, label, torch.ones_like(masks)))
        row = torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), label.view(-1), label_smoothing=0.1, reduction="none")
        out["loss_rows"] = row.numpy().astype(np.float32)
        # ---- full forward: image -> tokenizer -> masking -> transformer -> loss ----
        img, noise_f = train_forward_inputs()
        with inject_rand(noise_f):
            loss = pipe(img, text=["a photo"], mask_ratio=0.75)
        out["fwd_loss"] = np.float32(loss)
        with inject_rand(noise_f):
            loss_nt = pipe(img, text=None, mask_ratio=0.5)
        out["fwd_loss_notext"] = np.float32(loss_nt)
        _, ids, _ = pipe.to_latent(img)
        out["fwd_ids"] = ids.numpy().astype(np.int16)
    out["x_sum"] = np.float64(x.double().sum())
    out["noise_sum"] = np.float64(noise.double().sum())
    out["logits_sum"] = np.float64(logits.double().sum())
    np.savez_compressed(GOLD / "stage2_train.npz", **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
