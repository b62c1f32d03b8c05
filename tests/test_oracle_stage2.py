"""Numpy oracle of the stage-2 path against the reference's golden outputs (CPU)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import paintmind_oracle as O
from paintmind_b200.config import ver2cfg
from paintmind_b200.utils import synthetic
from stage2_inputs import TINY2, full_step_inputs, tiny_inputs


def test_oracle_cond_transformer_tiny():
    g = load_golden("stage2_tiny.npz")
    cfg1 = ver2cfg["vit-tiny-test"]
    for name, ctx_dim in (("same", 128), ("proj", 96)):
        sd = {k: v.numpy() for k, v in synthetic.make_stage2_state_dict(TINY2, cfg1, seed=5, context_dim=ctx_dim).items()}
        tokens, context = tiny_inputs(ctx_dim)
        logits = O.cond_transformer_forward(tokens.numpy(), context.numpy(), sd, TINY2)
        np.testing.assert_allclose(logits, g[f"logits_{name}"], atol=2e-4, rtol=0)
        if name == "same":
            np.testing.assert_allclose(O.cond_transformer_forward(tokens.numpy(), None, sd, TINY2), g["logits_nocontext"], atol=2e-4, rtol=0)


@pytest.mark.parametrize("version,fixture", [("paintmindv1", "stage2_step.npz"), ("paintmindv1-128", "stage2_step_256.npz")])
def test_oracle_maskgit_step_full_size(version, fixture):
    """1024 tokens (the reference's registered pipeline) and the labelled 256-token variant (image_size 128)."""
    g = load_golden(fixture)
    cfg2 = ver2cfg[version]
    cfg1 = ver2cfg[cfg2["stage1"]]
    N = (cfg1["enc"]["image_size"] // cfg1["enc"]["patch_size"]) ** 2
    sd = {("vqgan." + k): v.numpy() for k, v in synthetic.make_vqgan_state_dict(cfg1, seed=0).items()}
    sd.update({k: v.numpy() for k, v in synthetic.make_stage2_state_dict(cfg2, cfg1, seed=1, context_dim=1024).items()})
    text, ids, u = full_step_inputs(N=N)
    assert abs(float(text.double().sum()) - float(g["text_sum"])) < 1e-6 and abs(float(u.double().sum()) - float(g["u_sum"])) < 1e-3
    np.testing.assert_array_equal(ids.numpy(), g["ids_in"].astype(np.int64))
    new_ids, img, pred, logits, scores, k = O.sample_step(ids.numpy(), float(g["mask_ratio"]), text.numpy(), 5, 0.75, u.numpy(), sd, cfg2, cfg1)
    assert k == int(g["k"])
    np.testing.assert_allclose(logits[0, ::16, ::16], g["logits_sub"], atol=3e-4, rtol=0)
    ref_pred = g["pred_ids"].astype(np.int64)
    agree = (pred[0] == ref_pred).mean()
    assert agree > 0.995, agree          # fp32 summation-order noise can flip a near-tie in the top-5 / gumbel arg-max
    same = pred[0] == ref_pred
    np.testing.assert_allclose(scores[0][same], g["scores"][same], atol=1e-5, rtol=0)
    ref_new = g["new_ids"].astype(np.int64)
    assert (new_ids[0] == 8192).sum() == (ref_new == 8192).sum() == k
    assert ((new_ids[0] == 8192) != (ref_new == 8192)).sum() <= 8
    np.testing.assert_allclose(O.ids2tokens(ids.numpy(), sd["vqgan.quantize.embedding.weight"], sd["mask_token"])[0, :8], g["tokens_head"], atol=0, rtol=0)
    if same.all():
        np.testing.assert_allclose(img[0, :, ::4, ::4], g["img_sub"], atol=1e-3, rtol=0)


def test_schedule_matches_reference_probe():
    # SURVEY.md §3.3 [probe]: k per step at T = 12
    ks = [k for _, k, _ in O.generate_schedule(12, 1.0, 1024)]
    assert ks == [1015, 989, 946, 886, 812, 724, 623, 512, 391, 265, 133, 1]
