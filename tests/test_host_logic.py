"""Host-side logic that needs no GPU: config / factory surface, state_dict contract, weight repacking algebra."""
import numpy as np
import pytest
import torch

import paintmind_b200 as pm
from paintmind_b200.config import Config, ver2cfg
from paintmind_b200.engine import fold_layernorm, pack_swiglu_w12, pack_w3
from paintmind_b200.modules.mlp import swiglu_hidden
from paintmind_b200.utils import synthetic


def test_config_roundtrip(tmp_path):
    c = Config(ver2cfg["vit-s-vqgan"])
    assert c.n_embed == 8192 and c.embed_dim == 32 and c.beta == 0.25
    assert c.enc["dim"] == 512 and c.enc["depth"] == 8 and c.dec["out_channels"] == 3
    p = tmp_path / "c.json"
    c.to_json(p)
    d = Config()
    d.from_json(p)
    assert d.to_dict() == c.to_dict()
    assert swiglu_hidden(2048) == 1368 and swiglu_hidden(4096) == 2736          # modules/mlp.py:53


def test_factory_errors_match_reference():
    with pytest.raises(KeyError):
        pm.create_model(arch="vqgan", version="nope", pretrained=False)          # ver2cfg[version] (factory.py:7)
    with pytest.raises(ValueError, match="failed to load arch named foo"):
        pm.create_model(arch="foo", version="vit-s-vqgan", pretrained=False)     # factory.py:14


def test_vqmodel_state_dict_contract():
    """222 tensors, same keys/shapes as the reference VQModel; strict load works."""
    m = pm.create_model(arch="vqgan", version="vit-s-vqgan", pretrained=False)
    sd = m.state_dict()
    assert len(sd) == 222 and sum(v.numel() for v in sd.values()) == 52_032_992
    assert sd["encoder.to_patch_embedding.0.weight"].shape == (512, 3, 8, 8)
    assert sd["encoder.transformer.layers.7.ffnet.w12.weight"].shape == (2736, 512)
    assert sd["decoder.transformer.layers.0.attn1.to_out.0.bias"].shape == (512,)
    assert sd["decoder.proj.weight"].shape == (192, 512) and sd["quantize.embedding.weight"].shape == (8192, 32)
    assert "encoder.to_patch_embedding.0.bias" not in sd and "encoder.transformer.layers.0.attn1.to_q.bias" not in sd
    syn = synthetic.make_vqgan_state_dict(ver2cfg["vit-s-vqgan"], seed=0)
    assert set(syn) == set(sd) and all(syn[k].shape == sd[k].shape for k in sd)
    res = m.load_state_dict(syn, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_cpu_model_fails_loudly():
    """No CPU fallback: the product path raises instead of computing on the host."""
    m = pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=False).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m.encode(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.quantize(torch.zeros(1, 4, 32))


def test_fold_layernorm_identity():
    g = torch.Generator().manual_seed(0)
    K, N, M = 64, 48, 33
    W, b = torch.randn(N, K, generator=g).double(), torch.randn(N, generator=g).double()
    gamma, beta = (torch.rand(K, generator=g) + 0.5).double(), torch.randn(K, generator=g).double()
    x = torch.randn(M, K, generator=g).double() * 3 + 1
    Wf, colsum, bias = fold_layernorm(W, b, gamma, beta)
    mu = x.mean(1, keepdim=True)
    rstd = (x.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
    folded = rstd * (x @ Wf.double().t() - mu * colsum.double()[None]) + bias.double()[None]
    ref = torch.nn.functional.layer_norm(x, (K,), gamma, beta, 1e-5) @ W.t() + b
    # only the bf16 rounding of W' separates the two
    ref_bf = ((x - mu) * rstd) @ Wf.double().t() + (W @ beta + b)[None]
    torch.testing.assert_close(folded, ref_bf, atol=2e-5, rtol=0)   # colsum / bias are kept in fp32
    assert (folded - ref).abs().max() < 0.15


def test_swiglu_repack_layout_and_padding():
    g = torch.Generator().manual_seed(1)
    K, h = 32, 200                                   # hidden 200 -> padded 256 = 2 tiles
    w12, b12 = torch.randn(2 * h, K, generator=g), torch.randn(2 * h, generator=g)
    ones, zeros = torch.ones(K), torch.zeros(K)
    Wp, cs, bp, hp = pack_swiglu_w12(w12, b12, ones, zeros)
    assert hp == 256 and Wp.shape == (512, K) and bp.shape == (512,)
    for t in range(2):
        lo, hi = t * 128, min((t + 1) * 128, h)
        n = hi - lo
        torch.testing.assert_close(Wp[t * 256:t * 256 + n].float(), w12[lo:hi].bfloat16().float())            # gate rows
        torch.testing.assert_close(Wp[t * 256 + 128:t * 256 + 128 + n].float(), w12[h + lo:h + hi].bfloat16().float())  # value rows
        torch.testing.assert_close(bp[t * 256:t * 256 + n], b12[lo:hi])
        torch.testing.assert_close(bp[t * 256 + 128:t * 256 + 128 + n], b12[h + lo:h + hi])
    assert float(Wp[256 + 72:256 + 128].abs().max()) == 0 and float(bp[256 + 72:256 + 128].abs().max()) == 0   # padding
    w3 = torch.randn(16, h, generator=g)
    w3p = pack_w3(w3, hp)
    assert w3p.shape == (16, 256) and float(w3p[:, h:].abs().max()) == 0
    # emulate the kernel epilogue on the packed layout and compare with the reference SwiGLU (mlp.py:27-31)
    x = torch.randn(5, K, generator=g)
    acc = x.bfloat16().float() @ Wp.float().t() + bp
    hid = torch.cat([torch.nn.functional.silu(acc[:, t * 256:t * 256 + 128]) * acc[:, t * 256 + 128:(t + 1) * 256] for t in range(2)], 1)
    out = hid @ w3p.float().t()
    x12 = x.bfloat16().float() @ w12.bfloat16().float().t() + b12
    ref = (torch.nn.functional.silu(x12[:, :h]) * x12[:, h:]) @ w3.bfloat16().float().t()
    torch.testing.assert_close(out, ref, atol=1e-4, rtol=1e-4)


def test_pipeline_surface_and_schedule():
    from paintmind_b200.generate import mask_schedule
    ks = [max(int(mask_schedule((s + 1) / 12) * 1024), 1) for s in range(12)]
    assert ks == [1015, 989, 946, 886, 812, 724, 623, 512, 391, 265, 133, 1]      # SURVEY.md §3.3 probe
    assert isinstance(mask_schedule(0.5), np.floating)


def test_256_token_variant_is_registered_under_its_own_name():
    """a 256-token number must come from an explicitly registered image_size-128 configuration; the reference's two
    names keep the reference's values (1024 tokens)."""
    from paintmind_b200.config import ver2cfg
    v1, v128 = ver2cfg["vit-s-vqgan"], ver2cfg["vit-s-vqgan-128"]
    assert v1["enc"]["image_size"] == v1["dec"]["image_size"] == 256
    assert v128["enc"]["image_size"] == v128["dec"]["image_size"] == 128
    assert {k: v for k, v in v128["enc"].items() if k != "image_size"} == {k: v for k, v in v1["enc"].items() if k != "image_size"}
    assert (v128["enc"]["image_size"] // v128["enc"]["patch_size"]) ** 2 == 256
    p1, p128 = ver2cfg["paintmindv1"], ver2cfg["paintmindv1-128"]
    assert p1["stage1"] == "vit-s-vqgan" and p128["stage1"] == "vit-s-vqgan-128"
    assert {k: v for k, v in p128.items() if k != "stage1"} == {k: v for k, v in p1.items() if k != "stage1"}


def test_training_forward_has_no_cpu_fallback():
    """VQModel.forward under autograd (the generator training step) on a CPU model must fail loudly, like the inference path."""
    import pytest
    import torch
    import paintmind_b200 as pm
    model = pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=False).train()
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(x)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        model(x)


def test_allreduce_gradients_is_a_noop_without_a_process_group():
    import torch
    from paintmind_b200 import dist as pmdist
    lin = torch.nn.Linear(3, 2)
    lin.weight.grad = torch.ones_like(lin.weight)
    assert pmdist.allreduce_gradients(lin) is None and torch.equal(lin.weight.grad, torch.ones(2, 3))


def test_from_pretrained_round_trip_through_the_factory(tmp_path):
    """create_model(..., pretrained=True, checkpoint_path=...) (reference factory.py:6-21, vqmodel.py:43-44,
    generate.py:73-75): a reference-layout .pt written by torch.save(model.state_dict()) loads through the factory,
    strictly, onto whatever device the parameters live on."""
    import paintmind_b200 as pm
    from paintmind_b200.config import ver2cfg
    from paintmind_b200.utils import synthetic
    cfg = ver2cfg["vit-tiny-test"]
    sd = synthetic.make_vqgan_state_dict(cfg, seed=3)
    path = tmp_path / "vit-tiny-test.pt"
    torch.save(sd, path)
    m = pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=True, checkpoint_path=str(path))
    got = m.state_dict()
    assert set(got.keys()) == set(sd.keys())
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    # a key the architecture does not have is an error (strict load, as the reference)
    bad = dict(sd)
    bad["decoder.extra.weight"] = torch.zeros(1)
    torch.save(bad, path)
    with pytest.raises(RuntimeError):
        pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=True, checkpoint_path=str(path))
    # no checkpoint and no network: explicit failure, not a silent random init
    with pytest.raises(RuntimeError):
        pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=True)


def test_pipeline_checkpoint_with_t5_keys_loads(tmp_path):
    """The reference's Pipeline registers the frozen T5 encoder as a submodule, so RootYuan/paintmindv1.pt carries
    `text_model.transformer.*` tensors.  The text encoder is a pluggable callable here: those keys are dropped, every
    other key is checked strictly."""
    import paintmind_b200 as pm
    from paintmind_b200.config import ver2cfg
    from paintmind_b200.utils import synthetic
    ver2cfg.setdefault("pipeline-tiny-test", dict(ver2cfg["paintmindv1"], stage1="vit-tiny-test", dim=64, dim_head=64, mlp_dim=128,
                                                 num_head=1, depth=1))
    try:
        pipe = pm.create_model(arch="pipeline", version="pipeline-tiny-test", pretrained=False)
        sd = {k: v.clone() for k, v in pipe.state_dict().items()}
        sd["text_model.transformer.shared.weight"] = torch.zeros(4, 4)
        sd["text_model.transformer.encoder.block.0.layer.0.SelfAttention.q.weight"] = torch.zeros(4, 4)
        path = tmp_path / "pipe.pt"
        torch.save(sd, path)
        pipe2 = pm.create_model(arch="pipeline", version="pipeline-tiny-test", pretrained=True, checkpoint_path=str(path))
        assert all(torch.equal(v, sd[k]) for k, v in pipe2.state_dict().items())
        sd["transformer.bogus"] = torch.zeros(1)
        torch.save(sd, path)
        with pytest.raises(RuntimeError):
            pm.create_model(arch="pipeline", version="pipeline-tiny-test", pretrained=True, checkpoint_path=str(path))
    finally:
        ver2cfg.pop("pipeline-tiny-test", None)


def test_model_copies_build_their_own_engine():
    """copy.deepcopy / pickle of a VQModel must not share (or try to serialise) the engine of the original."""
    import copy
    import pickle
    import paintmind_b200 as pm
    m = pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=False)
    m.engine()
    c = copy.deepcopy(m)
    assert c._engine is None and c.engine() is not m.engine() and c.engine().model is c
    r = pickle.loads(pickle.dumps(m))
    assert r._engine is None and all(torch.equal(a, b) for a, b in zip(r.state_dict().values(), m.state_dict().values()))


def test_u8_normalise_formula_is_exact():
    """patchify8_u8 (csrc/pm_rowops.cu::norm_u8_fast) replaces the reference's ingest chain — T.ToTensor() = u / 255, then
    T.Normalize(0.5, 0.5) = (t - 0.5) / 0.5, both fp32 (reference utils/transform.py:17-18) — by one fused multiply-add on
    0x4B0000uu = 8388608 + u.  The two must agree after bf16 rounding for every one of the 256 byte values."""
    import numpy as np
    import torch
    u = np.arange(256, dtype=np.float32)
    t = (u / np.float32(255.0)).astype(np.float32)
    ref = torch.from_numpy(((t - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)).bfloat16()
    c = np.float32(0.007843137718737125)
    assert c == np.float32(2.0 / 255.0)
    k = np.float64(8388608.0) * np.float64(c) + 1.0
    assert k == 65794.0078125 and np.float32(k) == k            # the offset the FMA removes is representable
    magic = (np.uint32(0x4B000000) | np.arange(256, dtype=np.uint32)).view(np.float32)
    assert np.array_equal(magic, np.float32(8388608.0) + u)
    # the FMA is exact before its single rounding: float64 holds (8388608 + u) * c - k without error (24 + 24 bits)
    fast = torch.from_numpy((magic.astype(np.float64) * np.float64(c) - k).astype(np.float32)).bfloat16()
    assert torch.equal(fast.view(torch.int16), ref.view(torch.int16))


def test_nvtx_ranges_are_opt_in(monkeypatch):
    """ops.nvtx_range / nvtx_phase (tracing) do nothing unless PM_NVTX=1 and keep results and exceptions intact."""
    from paintmind_b200 import ops
    assert ops.NVTX is False or isinstance(ops.NVTX, bool)
    calls = []
    monkeypatch.setattr(ops, "NVTX", False)

    @ops.nvtx_phase("pm.test")
    def f(a, b=2):
        calls.append((a, b))
        return a + b
    assert f(1, b=3) == 4 and calls == [(1, 3)]
    with ops.nvtx_range("pm.test") as r:
        assert r.name == "pm.test"

    @ops.nvtx_phase("pm.raises")
    def g():
        raise ValueError("x")
    import pytest
    with pytest.raises(ValueError):
        g()
