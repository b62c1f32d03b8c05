"""The numpy oracle (oracle/paintmind_oracle.py) against fixtures produced by the REAL reference
(tests/make_golden.py).  This is what pins the oracle (the reference ships no tests of its own)."""
import numpy as np
import torch

from conftest import check_weight_checksums, load_golden, seeded_vqgan
from oracle import paintmind_oracle as O
from paintmind_b200.utils import synthetic


def _run_stage1(gold_name):
    g = load_golden(gold_name)
    cfg_name, batch, seed = str(g["cfg_name"]), int(g["batch"]), int(g["seed"])
    cfg, sd_t, sd = seeded_vqgan(cfg_name, seed)
    check_weight_checksums(g, sd_t)
    x = synthetic.make_images(batch, cfg["enc"]["image_size"], seed=seed + 100)
    assert abs(float(x.double().sum()) - float(g["x_sum"])) < 1e-6
    x = x.numpy()
    z_pre = O.vqmodel_latent(x, sd, cfg)
    np.testing.assert_allclose(z_pre, g["z_pre"], atol=2e-4, rtol=0)
    z_q, loss, idx = O.vq_forward(z_pre, sd["quantize.embedding.weight"], cfg["beta"])
    ref_idx = g["idx"].astype(np.int64)
    mism = idx != ref_idx
    # fp32 summation order differs between numpy and torch: indices may only differ at near-ties
    assert mism.mean() <= 0.002, f"{mism.sum()} index mismatches"
    assert np.all(g["gap"][mism] < 1e-4)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-4)
    np.testing.assert_allclose(z_q[~mism], g["z_q"][~mism], atol=2e-6, rtol=0)
    # decode from the REFERENCE's z_q so that decoder parity is independent of index near-ties
    s = int(g["rec_stride"])
    pre = O.vqmodel_decode(g["z_q"], sd, cfg, clamp=False)
    np.testing.assert_allclose(pre[:, :, ::s, ::s], g["pre_sub"], atol=5e-4, rtol=0)
    rec = np.clip(pre, -1, 1)
    np.testing.assert_allclose(rec[:, :, ::s, ::s], g["rec_sub"], atol=5e-4, rtol=0)
    np.testing.assert_allclose(rec.mean(dtype=np.float64), float(g["rec_mean"]), atol=1e-5)


def test_oracle_stage1_tiny():
    _run_stage1("stage1_tiny.npz")


def test_oracle_stage1_vit_s():
    _run_stage1("stage1_vit_s.npz")


def test_oracle_vq_microbench():
    g = load_golden("vq_microbench.npz")
    gen = torch.Generator().manual_seed(0)
    z = torch.nn.functional.normalize(torch.randn(65536, 32, generator=gen), dim=-1)
    E = torch.randn(8192, 32, generator=gen)
    assert abs(float(z.double().sum()) - float(g["z_sum"])) < 1e-6
    assert abs(float(E.double().sum()) - float(g["E_sum"])) < 1e-6
    z_q, loss, idx = O.vq_forward(z.numpy(), E.numpy(), 0.25)
    ref_idx = g["idx"].astype(np.int64)
    mism = idx != ref_idx
    assert mism.sum() <= 4 and np.all(g["gap"][mism] < 1e-5)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-5)
    np.testing.assert_allclose(z_q[:64], g["z_q_head"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(O.vq_decode_from_indice(ref_idx[:16], E.numpy()), g["dec_head"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(O.vq_top2_gap(z.numpy()[:2048], E.numpy()), g["gap"][:2048], atol=2e-6, rtol=0)


# ---- the torch-CPU flavour of the oracle (the timed CPU arm of bench.py) against the same fixtures ----
def _run_stage1_torch(gold_name):
    from oracle import paintmind_oracle_torch as OT
    g = load_golden(gold_name)
    cfg_name, batch, seed = str(g["cfg_name"]), int(g["batch"]), int(g["seed"])
    cfg, sd_t, _ = seeded_vqgan(cfg_name, seed)
    check_weight_checksums(g, sd_t)
    x = synthetic.make_images(batch, cfg["enc"]["image_size"], seed=seed + 100)
    with torch.no_grad():
        z_pre = OT.vqmodel_latent(x, sd_t, cfg)
        np.testing.assert_allclose(z_pre.numpy(), g["z_pre"], atol=2e-4, rtol=0)
        z_q, loss, idx = OT.vq_forward(z_pre, sd_t["quantize.embedding.weight"], cfg["beta"])
        mism = idx.numpy() != g["idx"].astype(np.int64)
        assert mism.mean() <= 0.002 and np.all(g["gap"][mism] < 1e-4)
        np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-4)
        s = int(g["rec_stride"])
        rec = OT.vqmodel_decode(torch.from_numpy(g["z_q"]), sd_t, cfg).numpy()
    np.testing.assert_allclose(rec[:, :, ::s, ::s], g["rec_sub"], atol=5e-4, rtol=0)
    np.testing.assert_allclose(rec.mean(dtype=np.float64), float(g["rec_mean"]), atol=1e-5)


def test_torch_oracle_stage1_tiny():
    _run_stage1_torch("stage1_tiny.npz")


def test_torch_oracle_stage1_vit_s():
    _run_stage1_torch("stage1_vit_s.npz")


def test_oracle_headline_fixture_sample():
    """Two of the 256 images of the batch-256 headline fixture (tests/make_golden_headline.py) through the numpy oracle:
    pins the fixture the GPU test at the BASELINE configs[2] size compares against (tests/test_gpu_headline.py)."""
    g = load_golden("headline_vit_s_b256.npz")
    B, seed, img_seed = int(g["batch"]), int(g["seed"]), int(g["img_seed"])
    ts, ps = int(g["tok_stride"]), int(g["pix_stride"])
    cfg, sd_t, sd = seeded_vqgan("vit-s-vqgan", seed)
    check_weight_checksums(g, sd_t)
    x = synthetic.make_images(B, 256, seed=img_seed)
    assert abs(float(x.double().sum()) - float(g["x_sum"])) < 1e-6 * max(1.0, abs(float(g["x_sum"])))
    pick = [0, B - 1]
    z_pre = O.vqmodel_latent(x[pick].numpy(), sd, cfg)
    np.testing.assert_allclose(z_pre[:, ::ts], g["z_pre_sub"][pick], atol=2e-4, rtol=0)
    z_q, loss, idx = O.vq_forward(z_pre, sd["quantize.embedding.weight"], cfg["beta"])
    ref_idx = g["idx"][pick].astype(np.int64)
    mism = idx != ref_idx
    assert mism.mean() <= 0.002 and np.all(g["gap"][pick][mism].astype(np.float32) < 1e-4)
    gap = O.vq_top2_gap(z_pre.reshape(-1, 32), sd["quantize.embedding.weight"]).reshape(2, -1)
    np.testing.assert_allclose(gap, g["gap"][pick].astype(np.float32), atol=3e-6, rtol=1e-3)      # fp16 storage, rounded up
    zq_ref = O.vq_decode_from_indice(ref_idx.reshape(-1), sd["quantize.embedding.weight"]).reshape(2, -1, 32)
    rec = O.vqmodel_decode(zq_ref, sd, cfg)
    np.testing.assert_allclose(rec[:, :, ::ps, ::ps], g["rec_sub"][pick].astype(np.float32), atol=1.5e-3, rtol=0)   # fp16 sample
    np.testing.assert_allclose(rec.mean(axis=(1, 2, 3), dtype=np.float64), g["rec_mean"][pick], atol=2e-5)
