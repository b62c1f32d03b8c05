"""Numpy oracle of the stage-2 TRAINING forward (generate.py:78-146) against fixtures produced
by the unmodified reference (tests/make_golden_train.py)."""
import numpy as np

from conftest import load_golden
from oracle import paintmind_oracle as O
from paintmind_b200.config import ver2cfg
from paintmind_b200.utils import synthetic
from stage2_inputs import loss_inputs, masking_inputs


def _mask_token():
    return synthetic.make_stage2_state_dict(ver2cfg["paintmindv1"], ver2cfg["vit-s-vqgan"], seed=1, context_dim=1024)["mask_token"].numpy()


def test_oracle_random_masking():
    g = load_golden("stage2_train.npz")
    x, noise = masking_inputs()
    assert abs(float(x.double().sum()) - float(g["x_sum"])) < 1e-9 and abs(float(noise.double().sum()) - float(g["noise_sum"])) < 1e-9
    mt = _mask_token()
    for tag, ratio in (("75", 0.75), ("30", 0.3), ("tiny", 0.0001)):
        xm, mask = O.random_masking(x.numpy(), ratio, mt, noise.numpy())
        ref_mask = np.unpackbits(g[f"mask_{tag}"], axis=1)[:, :1024]
        np.testing.assert_array_equal(mask.astype(np.uint8), ref_mask)
        assert int(mask[0].sum()) == max(int(1024 * ratio), 1)
        np.testing.assert_array_equal(xm[:, :16], g[f"xm_{tag}_head"])
        assert abs(float(xm.astype(np.float64).sum()) - float(g[f"xm_{tag}_sum"])) < 1e-6


def test_oracle_masked_ce_loss():
    g = load_golden("stage2_train.npz")
    logits, label, masks = loss_inputs()
    assert abs(float(logits.double().sum()) - float(g["logits_sum"])) < 1e-6
    rows = O.ce_label_smooth_rows(logits.numpy().reshape(-1, 8192), label.numpy().reshape(-1))
    np.testing.assert_allclose(rows, g["loss_rows"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(O.masked_ce_loss(logits.numpy(), label.numpy(), masks.numpy()), g["loss_unit"], rtol=1e-6)
    np.testing.assert_allclose(O.masked_ce_loss(logits.numpy(), label.numpy(), np.ones_like(masks.numpy())), g["loss_unit_allmask"], rtol=1e-6)


def test_oracle_train_forward_full_size():
    import torch
    from stage2_inputs import train_forward_inputs
    g = load_golden("stage2_train.npz")
    cfg1, cfg2 = ver2cfg["vit-s-vqgan"], ver2cfg["paintmindv1"]
    sd = {("vqgan." + k): v.numpy() for k, v in synthetic.make_vqgan_state_dict(cfg1, seed=0).items()}
    sd.update({k: v.numpy() for k, v in synthetic.make_stage2_state_dict(cfg2, cfg1, seed=1, context_dim=1024).items()})
    img, noise = train_forward_inputs()
    loss, ids, mask, _ = O.pipeline_train_forward(img.numpy(), None, 0.5, noise.numpy(), sd, cfg2, cfg1)
    assert (ids != g["fwd_ids"].astype(np.int64)).mean() < 0.002
    assert int(mask.sum()) == 512
    np.testing.assert_allclose(loss, g["fwd_loss_notext"], rtol=1e-5)
    text = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(1234)).numpy()     # the T5 stand-in (seed 1234)
    loss, _, _, _ = O.pipeline_train_forward(img.numpy(), text, 0.75, noise.numpy(), sd, cfg2, cfg1)
    np.testing.assert_allclose(loss, g["fwd_loss"], rtol=1e-5)
