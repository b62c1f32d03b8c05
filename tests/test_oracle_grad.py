"""Backward path: the differentiable torch oracle (oracle/paintmind_oracle_torch.py under autograd)
against gradient fixtures produced by the unmodified reference under torch autograd (tests/make_golden_grad.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import check_weight_checksums, load_golden
from grad_sampling import SMALL, grad_sample_positions
from oracle import paintmind_oracle_torch as OT
from paintmind_b200.config import ver2cfg
from paintmind_b200.utils import synthetic


def objective(rec, closs, img):
    """The path-only part of the generator loss (utils/trainer.py:207-208,215)."""
    return closs + F.l1_loss(rec, img) + F.mse_loss(rec, img)


def fixture_grad(gold, i, name, numel):
    """(positions or None, values) of gradient tensor i as the fixture holds it."""
    v = gold[f"g{i}"]
    return (None, v) if numel <= SMALL else (grad_sample_positions(name, numel), v)


@pytest.mark.parametrize("cfg_name,fixture", [("vit-tiny-test", "stage1_grad_tiny.npz"), ("vit-s-vqgan", "stage1_grad_vit_s.npz")])
def test_oracle_gradients_match_reference(cfg_name, fixture):
    gold = load_golden(fixture)
    cfg = ver2cfg[cfg_name]
    sd = synthetic.make_vqgan_state_dict(cfg, seed=int(gold["seed"]))
    check_weight_checksums(gold, sd)
    img = synthetic.make_images(int(gold["batch"]), cfg["enc"]["image_size"], seed=int(gold["seed"]) + 200)
    assert abs(float(img.double().abs().sum()) - float(gold["img_abs_sum"])) < 1e-6
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rec, closs, idx = OT.vqmodel_forward_train(img, sdg, cfg)
    L = objective(rec, closs, img)
    L.backward()
    assert abs(float(L.detach()) - float(gold["loss"])) < 2e-5
    assert abs(float(closs.detach()) - float(gold["codebook_loss"])) < 2e-6
    np.testing.assert_array_equal(idx.numpy().astype(np.int32), gold["indices"])
    for i, n in enumerate(gold["names"].tolist()):
        g = sdg[n].grad.reshape(-1)
        pos, ref = fixture_grad(gold, i, n, g.numel())
        got = g.numpy() if pos is None else g.numpy()[pos]
        # fp32 on both sides; summation order differs (bmm vs einsum, functional vs module): 1e-3 of the tensor's scale
        scale = float(gold["norms"][i]) / np.sqrt(g.numel())
        assert np.abs(got - ref).max() <= 2e-3 * max(scale, 1e-12) + 1e-9, (n, np.abs(got - ref).max(), scale)
        assert abs(float(g.double().norm()) - float(gold["norms"][i])) <= 1e-3 * float(gold["norms"][i]) + 1e-12, n


def test_packed_w12_row_map_roundtrip():
    """train._BlockBwd: packed (tile-interleaved, padded) w12 rows <-> reference rows is a bijection on the real rows."""
    from types import SimpleNamespace

    from paintmind_b200.stage1.layers import Layer
    from paintmind_b200.train import _BlockBwd
    layer = Layer(dim=64, dim_head=64, mlp_dim=300, num_head=1)
    h = layer.ffnet.w12.weight.shape[0] // 2
    hp = (h + 127) // 128 * 128
    bw = _BlockBwd(layer, SimpleNamespace(hp=hp))
    assert sorted(bw.orig_rows.tolist()) == list(range(2 * h))
    w12p = bw.w_12_t.t().float()                                  # [2 hp, D] packed
    w = layer.ffnet.w12.weight.detach().to(torch.bfloat16).float()
    back = torch.empty_like(w)
    back[bw.orig_rows] = w12p[bw.packed_rows]
    assert torch.equal(back, w)
    pad = torch.ones(2 * hp, dtype=torch.bool)
    pad[bw.packed_rows] = False
    assert float(w12p[pad].abs().sum()) == 0.0
    # gate row j of tile T sits 128 rows before its value row
    T, j = 1, 5
    assert torch.equal(w12p[T * 256 + j], w[T * 128 + j]) and torch.equal(w12p[T * 256 + 128 + j], w[h + T * 128 + j])
