"""GPU parity of the generator BACKWARD path, through the C-ABI.

Kernel level: every backward kernel against torch autograd (fp32) of the reference formula on the same bf16-rounded
operands.  Model level: VQModel.forward(...).backward() against (a) torch autograd of the fp32 torch oracle on the GPU
with the code indices forced to the CUDA path's (near-tie flips excluded by construction), all 222 tensors in full, and
(b) the fixtures the unmodified reference produced under autograd (sampled entries + norms).

Stated tolerances: bf16 operands with fp32 accumulation -> relative L2 per gradient tensor <= 5 % against fp32 autograd
(measured on B200: median 2.2 %, max 3.1 %; the reference itself under bf16 autocast sits at median 3.3 %, max 10 % from its
own fp32 gradients — `autocast_rel_l2` in the fixtures)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import check_weight_checksums, load_golden
from grad_sampling import SMALL, grad_sample_positions

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def objective(rec, closs, img):
    return closs + F.l1_loss(rec.float(), img) + F.mse_loss(rec.float(), img)


@pytest.fixture(scope="module")
def ops(cuda_device):
    from paintmind_b200 import ops as _ops
    return _ops


def rnd(gen, *shape, scale=1.0):
    return torch.randn(*shape, device="cuda", generator=gen) * scale


@pytest.mark.parametrize("M,N,K", [(4096, 1536, 512), (2048, 512, 1408), (2048, 64, 512), (2048, 512, 64), (1000, 192, 512), (4160, 2816, 512)])
def test_wgrad(ops, M, N, K):
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    dy, x = rnd(gen, M, N).bfloat16(), rnd(gen, M, K).bfloat16()
    out = torch.full((N, K), 3.0, device="cuda")
    ops.wgrad(dy, x, out)
    ref = dy.float().t() @ x.float()
    assert rel_l2(out, ref) < 1e-5
    again = torch.empty_like(out)
    ops.wgrad(dy, x, again)
    assert torch.equal(out, again), "split-K reduction must be deterministic"
    ops.wgrad(dy, x, out, accumulate=True)
    assert rel_l2(out, 2 * ref) < 1e-5


def test_colsum(ops):
    gen = torch.Generator(device="cuda").manual_seed(3)
    for M, N in ((4096, 512), (300, 2816), (16, 1024 * 128)):
        x = rnd(gen, M, N).bfloat16()
        out = torch.empty(N, device="cuda")
        ops.colsum(x, out)
        assert rel_l2(out, x.float().sum(0)) < 1e-5


@pytest.mark.parametrize("M,D", [(4096, 512), (1000, 128), (512, 1024)])
def test_layernorm_bwd(ops, M, D):
    gen = torch.Generator(device="cuda").manual_seed(D)
    x, dn, dres = rnd(gen, M, D).bfloat16(), rnd(gen, M, D).bfloat16(), rnd(gen, M, D).bfloat16()
    gamma, beta = (1 + 0.1 * rnd(gen, D)).contiguous(), 0.1 * rnd(gen, D)
    xr, gr, br = x.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), gr, br, 1e-5).backward(dn.float())
    dx, dgb = torch.empty_like(x), torch.empty(2, D, device="cuda")
    ops.layernorm_bwd(dn, x, gamma, dx, dgb, dres=dres)
    assert rel_l2(dx, xr.grad + dres.float()) < 4e-3          # bf16 store
    assert rel_l2(dgb[0], gr.grad) < 1e-5 and rel_l2(dgb[1], br.grad) < 1e-5


def test_swiglu_bwd(ops):
    gen = torch.Generator(device="cuda").manual_seed(5)
    M, hp = 1024, 384
    x12, dh = rnd(gen, M, 2 * hp).bfloat16(), rnd(gen, M, hp).bfloat16()
    h, d12 = torch.empty(M, hp, device="cuda", dtype=torch.bfloat16), torch.empty(M, 2 * hp, device="cuda", dtype=torch.bfloat16)
    b12 = torch.empty(2 * hp, device="cuda")
    ops.swiglu_bwd(x12, dh, h, d12, b12)
    t = x12.float().view(M, hp // 128, 2, 128)
    g, v = t[:, :, 0].reshape(M, hp).clone().requires_grad_(True), t[:, :, 1].reshape(M, hp).clone().requires_grad_(True)
    hr = F.silu(g) * v                                          # modules/mlp.py:29-30
    hr.backward(dh.float())
    dref = torch.stack([g.grad.view(M, -1, 128), v.grad.view(M, -1, 128)], dim=2).reshape(M, 2 * hp)
    assert rel_l2(h, hr) < 4e-3 and rel_l2(d12, dref) < 4e-3
    assert rel_l2(b12, dref.sum(0)) < 1e-5                      # bias gradient of w12 (packed order), summed in fp32 before rounding
    d12b = torch.empty_like(d12)
    ops.swiglu_bwd(x12, dh, None, d12b)                         # without h / without the sums
    assert torch.equal(d12, d12b)


@pytest.mark.parametrize("B,H,N", [(2, 8, 1024), (1, 2, 128), (2, 2, 64), (2, 3, 200)])
def test_attention_bwd(ops, B, H, N):
    gen = torch.Generator(device="cuda").manual_seed(N)
    inner, scale = H * 64, 0.125
    qkv = rnd(gen, B, N, 3 * inner).bfloat16()
    q, k, v = qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
    o = torch.empty(B, N, inner, device="cuda", dtype=torch.bfloat16)
    o32 = torch.empty(B, N, inner, device="cuda")
    lse = ops.lse_buffer(B, H, N, "cuda")
    ops.attention_train(q, k, v, o, H, scale, lse, o32)
    heads = lambda t: t.float().view(B, N, H, 64).permute(0, 2, 1, 3)
    qr, kr, vr = [heads(t).contiguous().requires_grad_(True) for t in (q, k, v)]
    sim = (qr * scale) @ kr.transpose(-1, -2)                  # modules/attention.py:52-54
    oref = sim.softmax(dim=-1) @ vr                            # :55-57
    assert rel_l2(lse, torch.logsumexp(sim, dim=-1) * 1.4426950408889634) < 1e-5
    assert rel_l2(heads(o32), oref) < 4e-3 and rel_l2(heads(o), oref) < 6e-3
    do = rnd(gen, B, N, inner).bfloat16()
    oref.backward(heads(do))
    dqkv = torch.zeros_like(qkv)
    ops.attention_bwd(q, k, v, o32, do, lse, dqkv[..., :inner], dqkv[..., inner:2 * inner], dqkv[..., 2 * inner:], H, scale)
    for got, ref in ((dqkv[..., :inner], qr.grad), (dqkv[..., inner:2 * inner], kr.grad), (dqkv[..., 2 * inner:], vr.grad)):
        assert rel_l2(heads(got), ref) < 6e-3
    again = torch.zeros_like(qkv)
    ops.attention_bwd(q, k, v, o32, do, lse, again[..., :inner], again[..., inner:2 * inner], again[..., 2 * inner:], H, scale)
    assert torch.equal(dqkv, again), "attention backward must be deterministic (no atomics)"


def test_vq_bwd(ops):
    gen = torch.Generator(device="cuda").manual_seed(9)
    M, n_e, beta = 4096, 8192, 0.25
    z, E, d_out = rnd(gen, M, 32), rnd(gen, n_e, 32), rnd(gen, M, 32, scale=0.01)
    d_loss = torch.tensor([0.7], device="cuda")
    zr, Er = z.clone().requires_grad_(True), E.clone().requires_grad_(True)
    zn = F.normalize(zr, dim=-1)                                                   # quantize.py:19
    idx = torch.argmax(zn.detach() @ F.normalize(E, dim=-1).t(), dim=1)
    zq = F.normalize(Er[idx], dim=-1)                                              # :29-30
    loss = beta * torch.mean((zq.detach() - zn) ** 2) + torch.mean((zq - zn.detach()) ** 2)   # :33
    out = zn + (zq - zn).detach()                                                  # :36
    ((out * d_out).sum() + loss * d_loss[0]).backward()
    dz, dzs, dE = torch.empty(M, 32, device="cuda"), torch.empty(M, 64, device="cuda", dtype=torch.bfloat16), torch.zeros(n_e, 32, device="cuda")
    ops.vq_bwd(z, idx, E, d_out, d_loss, beta, dz=dz, dz_split=dzs, dE=dE)
    assert rel_l2(dz, zr.grad) < 1e-5 and rel_l2(dzs[:, :32].float() + dzs[:, 32:].float(), zr.grad) < 1e-4
    assert rel_l2(dE, Er.grad) < 1e-5


def test_unpatchify_bwd(ops):
    gen = torch.Generator(device="cuda").manual_seed(2)
    B = 3
    g, rec = rnd(gen, B, 3, 256, 256), rnd(gen, B, 3, 256, 256).clamp(-1, 1)
    out = torch.empty(B * 1024, 192, device="cuda", dtype=torch.bfloat16)
    ops.unpatchify8_bwd(g, rec, out)
    ref = ((rec.abs() < 1).float() * g).view(B, 3, 32, 8, 32, 8).permute(0, 2, 4, 1, 3, 5).reshape(B * 1024, 192)
    assert torch.equal(out, ref.bfloat16())


def _model_and_grads(cfg_name, seed, batch, keep=None):
    import paintmind_b200 as pm
    from paintmind_b200.config import ver2cfg
    from paintmind_b200.utils import synthetic
    cfg = ver2cfg[cfg_name]
    sd = synthetic.make_vqgan_state_dict(cfg, seed=seed)
    model = pm.create_model(arch="vqgan", version=cfg_name, pretrained=False)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    model.train_engine().keep_attention = keep       # None: decided from free memory; False: recompute attention in backward
    img = synthetic.make_images(batch, cfg["enc"]["image_size"], seed=seed + 200).cuda()
    rec, closs = model(img)
    assert rec.requires_grad and closs.requires_grad
    L = objective(rec, closs, img)
    L.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    assert all(g is not None and torch.isfinite(g).all() for g in grads.values())
    return cfg, sd, model, img, float(L.detach()), grads


@pytest.mark.parametrize("cfg_name,seed,batch,keep", [("vit-tiny-test", 7, 3, False), ("vit-s-vqgan", 0, 2, True), ("vit-tiny-test", 7, 3, True)])
def test_model_gradients_vs_oracle_autograd(cuda_device, cfg_name, seed, batch, keep):
    from oracle import paintmind_oracle_torch as OT
    cfg, sd, model, img, L, grads = _model_and_grads(cfg_name, seed, batch, keep)
    # the indices the training forward itself chose (the oracle must differentiate the same quantisation); the inference path
    # runs the same kernels and picks the same codes
    idx = model.train_engine().last_indices.clone()
    with torch.no_grad():
        _, _, idx_inf = model.encode(img)
    assert torch.equal(idx_inf, idx)
    sdg = {k: v.cuda().clone().requires_grad_(True) for k, v in sd.items()}
    rec_o, closs_o, _ = OT.vqmodel_forward_train(img, sdg, cfg, idx=idx)
    Lo = objective(rec_o, closs_o, img)
    Lo.backward()
    assert abs(L - float(Lo.detach())) < 2e-3 * abs(float(Lo.detach()))
    worst = max((rel_l2(grads[n], sdg[n].grad), n) for n in grads)
    assert worst[0] < 0.05, worst
    # inference path afterwards still works and a second step reproduces the deterministic gradients bit for bit
    model.zero_grad(set_to_none=True)
    rec, closs = model(img)
    objective(rec, closs, img).backward()
    for n, p in model.named_parameters():
        if n != "quantize.embedding.weight":                      # fp32 atomics (scatter-add of rows)
            assert torch.equal(p.grad, grads[n]), n


@pytest.mark.parametrize("cfg_name,fixture", [("vit-tiny-test", "stage1_grad_tiny.npz"), ("vit-s-vqgan", "stage1_grad_vit_s.npz")])
def test_model_gradients_vs_reference_fixture(cuda_device, cfg_name, fixture):
    gold = load_golden(fixture)
    cfg, sd, model, img, L, grads = _model_and_grads(cfg_name, int(gold["seed"]), int(gold["batch"]))
    check_weight_checksums(gold, sd)
    assert abs(L - float(gold["loss"])) < 3e-3 * abs(float(gold["loss"]))
    bad = []
    for i, n in enumerate(gold["names"].tolist()):
        g = grads[n].reshape(-1).float().cpu().numpy()
        ref = gold[f"g{i}"]
        got = g if g.size <= SMALL else g[grad_sample_positions(n, g.size)]
        err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
        nrm = abs(np.linalg.norm(g.astype(np.float64)) - float(gold["norms"][i])) / float(gold["norms"][i])
        # the fp32 reference picks a different code for ~1.5 % of the tokens (near-ties, see test_gpu_stage1): those tokens
        # enter the decoder with another code vector and touch other rows of the codebook gradient, so this comparison
        # carries that on top of the bf16 noise (the forced-index test above is the tight one: 5 %).  For scale, the
        # reference's own bf16-autocast gradients are up to 10-15 % away from its fp32 ones (autocast_rel_l2).
        tol = 0.35 if n == "quantize.embedding.weight" else 0.15
        if err > tol or nrm > 0.06:
            bad.append((n, err, nrm))
    assert not bad, bad


def test_image_gradient_is_refused_loudly(cuda_device):
    import paintmind_b200 as pm
    from paintmind_b200.utils import synthetic
    model = pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=False).cuda()
    img = synthetic.make_images(1, 64, seed=1).cuda().requires_grad_(True)
    with pytest.raises(NotImplementedError):
        model(img)


def test_frozen_model_takes_inference_path(cuda_device):
    import paintmind_b200 as pm
    from paintmind_b200.utils import synthetic
    model = pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=False).cuda()
    img = synthetic.make_images(2, 64, seed=1).cuda()
    rec_t, loss_t = model(img)
    model.freeze()
    rec, loss = model(img)
    assert not rec.requires_grad and rec_t.requires_grad
    with torch.no_grad():
        rec_i, loss_i = model(img)
    assert torch.equal(rec, rec_i) and torch.equal(loss, loss_i)          # frozen == no_grad: both the inference engine
    # the training forward runs the same kernels (plus saved tensors for the backward): bit-identical
    assert torch.equal(rec, rec_t.detach()) and torch.equal(loss, loss_t.detach())


def _vit_s_model():
    import paintmind_b200 as pm
    from paintmind_b200.config import ver2cfg
    from paintmind_b200.utils import synthetic
    cfg = ver2cfg["vit-s-vqgan"]
    model = pm.create_model(arch="vqgan", version="vit-s-vqgan", pretrained=False)
    model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=0), strict=True)
    return model.cuda().train()


def _grads_of(model, img, weight, scale=1.0):
    model.zero_grad(set_to_none=True)
    rec, closs = model(img)
    ((rec * weight).sum() * scale).backward()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def test_full_size_gradient_properties(cuda_device):
    """Size-independent properties of the backward path at the full token count (vit-s-vqgan, 1024 tokens, batch 32):
    (1) homogeneity: scaling the loss by 2 scales every gradient by exactly 2 (powers of two commute with every bf16 / fp32
        rounding on the path) — bit-exact, including the split-K weight gradients;
    (2) additivity over images: no kernel couples samples, so the gradient of a sum-reduced loss over a batch equals the sum
        of the gradients of its two halves up to fp32 summation order (weight gradients split the token dimension
        differently for different batch sizes)."""
    from paintmind_b200.utils import synthetic
    model = _vit_s_model()
    B = 32
    img = synthetic.make_images(B, 256, seed=21).cuda()
    w = torch.randn(B, 3, 256, 256, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 1e-3
    g1 = _grads_of(model, img, w)
    g2 = _grads_of(model, img, w, scale=2.0)
    assert set(g1) == set(g2) and "encoder.to_patch_embedding.0.weight" in g1 and "quantize.embedding.weight" not in g1
    for n in g1:
        assert torch.equal(g2[n], 2.0 * g1[n]), n
    ga = _grads_of(model, img[: B // 2], w[: B // 2])
    gb = _grads_of(model, img[B // 2:], w[B // 2:])
    worst = max((rel_l2(ga[n] + gb[n], g1[n]), n) for n in g1)
    assert worst[0] < 2e-4, worst
