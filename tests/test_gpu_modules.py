"""Unit parity of the module surface against the numpy oracle on the
same inputs, plus shape edge cases of the kernels behind them (ragged M, batch 1, odd batches)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import paintmind_oracle as O

pytestmark = pytest.mark.gpu


def _sd(mod):
    return {k: v.detach().cpu().numpy() for k, v in mod.state_dict().items()}


@pytest.mark.parametrize("B,N,L,heads,dim,ctx_dim", [(2, 256, None, 4, 256, None), (1, 100, 77, 2, 128, 192), (3, 64, 64, 8, 512, 512)])
def test_cross_attention_module_vs_oracle(cuda_device, B, N, L, heads, dim, ctx_dim):
    """modules/attention.py:43-59 — self (context=None) and cross attention, ragged query / key lengths."""
    from paintmind_b200.modules.attention import CrossAttention
    torch.manual_seed(0)
    m = CrossAttention(query_dim=dim, context_dim=ctx_dim, heads=heads, dim_head=64).eval()
    x = torch.randn(B, N, dim)
    ctx = torch.randn(B, L, ctx_dim) if L is not None else None
    ref = O.attention(x.numpy(), _sd(m), "", heads, None if ctx is None else ctx.numpy())
    got = m.to(cuda_device)(x.to(cuda_device), None if ctx is None else ctx.to(cuda_device))
    assert got.dtype == torch.float32 and got.shape == x.shape
    err = (got.cpu().numpy() - ref)
    assert np.abs(err).max() < 0.03 and np.abs(err).mean() < 0.004, (np.abs(err).max(), np.abs(err).mean())


@pytest.mark.parametrize("rows,dim,mlp_dim", [(300, 512, 2048), (128, 128, 256), (1000, 1024, 4096)])
def test_swiglu_ffn_module_vs_oracle(cuda_device, rows, dim, mlp_dim):
    """modules/mlp.py:27-31,51-53 — hidden = (int(mlp_dim*2/3)+7)//8*8 (1368, 176, 2736): padded tiles, ragged rows."""
    from paintmind_b200.modules.mlp import SwiGLUFFNFused
    torch.manual_seed(1)
    m = SwiGLUFFNFused(in_features=dim, hidden_features=mlp_dim).eval()
    assert m.w3.weight.shape[1] == O.swiglu_hidden(mlp_dim)
    x = torch.randn(rows, dim)
    ref = O.swiglu_ffn(x.numpy(), _sd(m), "")
    got = m.to(cuda_device)(x.to(cuda_device))
    err = np.abs(got.cpu().numpy() - ref)
    assert err.max() < 0.03 and err.mean() < 0.003, (err.max(), err.mean())


def test_vector_quantizer_surface(cuda_device):
    """quantize.py:18-44: any float dtype in, fp32 z_q / loss and int64 indices out, shape preserved, beta honoured."""
    from paintmind_b200.stage1.quantize import VectorQuantizer
    torch.manual_seed(2)
    vq = VectorQuantizer(1000, 32, beta=0.5).to(cuda_device)
    z = torch.randn(3, 7, 5, 32, device=cuda_device)
    for dt in (torch.float32, torch.bfloat16, torch.float64):
        z_q, loss, idx = vq(z.to(dt))
        assert z_q.dtype == torch.float32 and idx.dtype == torch.int64 and loss.dtype == torch.float32
        assert z_q.shape == z.shape and idx.shape == z.shape[:-1] and loss.shape == ()
    zq_o, loss_o, idx_o = O.vq_forward(z.cpu().numpy(), vq.embedding.weight.detach().cpu().numpy(), 0.5)
    z_q, loss, idx = vq(z)
    gap = O.vq_top2_gap(z.cpu().numpy(), vq.embedding.weight.detach().cpu().numpy()).reshape(idx_o.shape)
    mism = idx.cpu().numpy() != idx_o
    assert np.all(gap[mism] < 1e-5)
    np.testing.assert_allclose(loss.item(), float(loss_o), rtol=1e-5)
    np.testing.assert_allclose(z_q.cpu().numpy()[~mism], zq_o[~mism], atol=2e-6)
    dec = vq.decode_from_indice(idx)
    np.testing.assert_allclose(dec.cpu().numpy(), O.vq_decode_from_indice(idx.cpu().numpy(), vq.embedding.weight.detach().cpu().numpy()), atol=1e-6)
    # non-contiguous input view
    zt = torch.randn(32, 40, device=cuda_device).t()
    z_q2, _, idx2 = vq(zt)
    z_q3, _, idx3 = vq(zt.contiguous())
    assert torch.equal(idx2, idx3) and torch.equal(z_q2, z_q3)


@pytest.mark.parametrize("batch", [1, 3, 5])
def test_odd_batches_match_single_image_results(cuda_device, batch):
    """Images are independent units: a batch of B gives the same tokens / pixels as B batches of one
    (ragged M = B*64 in the 256-row CTA-pair GEMM tiles, attention items, VQ row tiles)."""
    import paintmind_b200 as pm
    from paintmind_b200.config import ver2cfg
    from paintmind_b200.utils import synthetic
    cfg = ver2cfg["vit-tiny-test"]
    model = pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=False)
    model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=3))
    model = model.to(cuda_device).eval()
    x = synthetic.make_images(batch, 64, seed=50 + batch).to(cuda_device)
    z, loss, idx = model.encode(x)
    rec = model.decode(z)
    for b in range(batch):
        zb, _, ib = model.encode(x[b:b + 1])
        assert torch.equal(ib, idx[b:b + 1])
        assert torch.equal(zb, z[b:b + 1])
        assert torch.equal(model.decode(zb), rec[b:b + 1])


def test_encoder_decoder_standalone_modules(cuda_device):
    """Encoder.forward / Decoder.forward are callable on their own (layers.py:106-112, 145-152)."""
    import paintmind_b200 as pm
    from paintmind_b200.config import ver2cfg
    from paintmind_b200.utils import synthetic
    cfg = ver2cfg["vit-tiny-test"]
    sd = synthetic.make_vqgan_state_dict(cfg, seed=3)
    model = pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=False)
    model.load_state_dict(sd)
    model = model.to(cuda_device).eval()
    x = synthetic.make_images(2, 64, seed=9)
    tok = model.encoder(x.to(cuda_device))
    sd_np = {k: v.numpy() for k, v in sd.items()}
    ref_tok = O.encoder_forward(x.numpy(), sd_np, cfg["enc"])
    assert tok.dtype == torch.float32 and tuple(tok.shape) == ref_tok.shape
    assert np.abs(tok.cpu().numpy() - ref_tok).max() < 0.08
    t_in = torch.randn(2, 64, 128)
    img = model.decoder(t_in.to(cuda_device))
    ref_img = np.clip(O.decoder_forward(t_in.numpy(), sd_np, cfg["dec"]), -1, 1)
    assert np.abs(img.cpu().numpy() - ref_img).max() < 0.06


def test_weight_update_repacks(cuda_device):
    """The bf16 operands are re-derived when a parameter changes (load_state_dict / in-place update)."""
    import paintmind_b200 as pm
    from paintmind_b200.config import ver2cfg
    from paintmind_b200.utils import synthetic
    cfg = ver2cfg["vit-tiny-test"]
    model = pm.create_model(arch="vqgan", version="vit-tiny-test", pretrained=False).to(cuda_device).eval()
    x = synthetic.make_images(1, 64, seed=1).to(cuda_device)
    model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=3))
    _, _, idx_a = model.encode(x)
    model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=4))
    _, _, idx_b = model.encode(x)
    model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=3))
    _, _, idx_c = model.encode(x)
    assert torch.equal(idx_a, idx_c) and not torch.equal(idx_a, idx_b)


def test_gemm_row_periodic_residual_table(cuda_device):
    """`res_mod`: the position-embedding table handed to the GEMM as a row-periodic bf16 residual (staged by TMA)
    gives the same result as the fp32 `pos` path up to the table's bf16 rounding."""
    from paintmind_b200 import ops
    g = torch.Generator().manual_seed(4)
    for M, N, K, T in [(1024, 512, 192, 256), (2048, 512, 64, 1024), (384, 256, 64, 128)]:
        a = (torch.randn(M, K, generator=g) * 0.5).to(cuda_device).bfloat16()
        w = (torch.randn(N, K, generator=g) * 0.1).to(cuda_device).bfloat16()
        bias = (torch.randn(N, generator=g) * 0.1).to(cuda_device)
        pos = (torch.randn(T, N, generator=g) * 0.02).to(cuda_device)
        o1 = torch.empty(M, N, device=cuda_device, dtype=torch.bfloat16)
        o2 = torch.empty_like(o1)
        ops.gemm(a, w, o1, bias=bias, pos=pos)
        ops.gemm(a, w, o2, bias=bias, res=pos.bfloat16().contiguous(), res_mod=T)
        ref = a.float() @ w.float().t() + bias + pos.bfloat16().float().repeat(M // T, 1)
        assert (o2.float() - ref).abs().max() < 0.02 * max(1.0, ref.abs().max().item())
        assert (o1.float() - o2.float()).abs().max() < 0.02
        # statistics of the output rows still come out right on this path
        st = torch.empty(M, ops.stats_parts(N), 2, device=cuda_device)
        ops.gemm(a, w, o2, bias=bias, res=pos.bfloat16().contiguous(), res_mod=T, stats_out=st)
        # (the statistics are taken before the bf16 rounding of the stored values: N roundings of ~2^-9 |x| apart)
        torch.testing.assert_close(st[..., 0].sum(1), o2.float().sum(1), atol=0.3, rtol=1e-2)
    with pytest.raises(RuntimeError):
        ops.gemm(a, w, o2, res=pos.bfloat16().contiguous(), res_mod=100)       # not a multiple of the 128-row tile
