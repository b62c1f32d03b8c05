"""GPU parity of the uint8 pixel ingest / egress through the C-ABI.

Ingest = the reference's transform tail (utils/transform.py:17-18: ToTensor = u/255, Normalize(0.5, 0.5) =
(t-0.5)/0.5, fp32) fused into the patch extraction; egress = `restore` (reconstruct.py:11-16: (x+1)*0.5, HWC,
uint8(255*x) truncating) fused into the un-patchify epilogue.  Both are integer / exactly-rounded fp32
pipelines, so the bar is BIT-EXACT against the reference formulas evaluated by torch on the same data."""
import numpy as np
import pytest
import torch

from conftest import seeded_vqgan
from paintmind_b200 import ops

pytestmark = pytest.mark.gpu


def _reference_ingest(u8_nhwc):
    """ToTensor + Normalize(mean 0.5, std 0.5) exactly as torchvision evaluates them: fp32 ON THE HOST (the
    transform runs in the data loader; torch's CPU `div` is a true division, whereas its CUDA kernel multiplies by
    the rounded reciprocal of a scalar divisor and differs in the last bit for some byte values)."""
    dev = u8_nhwc.device
    t = u8_nhwc.cpu().permute(0, 3, 1, 2).to(torch.float32).div(255)
    return t.sub(0.5).div(0.5).contiguous().to(dev)


def _reference_restore(x_nchw):
    """reconstruct.py:11-16 on a batch."""
    x = (x_nchw + 1) * 0.5
    x = x.permute(0, 2, 3, 1).cpu().numpy()
    return (255 * x).astype(np.uint8)


def _model(cfg_name, sd, dev):
    import paintmind_b200 as pm
    m = pm.create_model(arch="vqgan", version=cfg_name, pretrained=False)
    m.load_state_dict(sd, strict=True)
    return m.to(dev).eval()


@pytest.mark.parametrize("B,C,H,W", [(2, 3, 256, 256), (3, 3, 64, 32), (1, 3, 8, 8), (5, 3, 24, 256), (2, 1, 32, 32), (1, 3, 16, 512)])
def test_patchify_fp32_vs_unfold(cuda_device, B, C, H, W):
    """pm_patchify8 (stage1/layers.py:82-83: the stride-8 patch conv's im2col, K order c, kh, kw) against torch unfold: the staged
    kernel (C = 3, W <= 256; ragged last block iterations, non-square images) and the direct kernel (other shapes)."""
    g = torch.Generator().manual_seed(11)
    img = (torch.rand(B, C, H, W, generator=g) * 2 - 1).to(cuda_device)
    M = B * (H // 8) * (W // 8)
    got = torch.full((M, C * 64), float("nan"), device=cuda_device, dtype=torch.bfloat16)
    ops.patchify8(img, got)
    torch.cuda.synchronize()
    ref = torch.nn.functional.unfold(img, kernel_size=8, stride=8).transpose(1, 2).reshape(M, C * 64).to(torch.bfloat16)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))


@pytest.mark.parametrize("B,S", [(1, 64), (3, 256), (2, 8)])
def test_patchify_u8_bit_exact(cuda_device, B, S):
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8)
    # every byte value occurs
    u8.view(-1)[:256] = torch.arange(256, dtype=torch.uint8)
    u8d = u8.to(cuda_device)
    M = B * (S // 8) ** 2
    got = torch.empty(M, 192, device=cuda_device, dtype=torch.bfloat16)
    want = torch.empty_like(got)
    ops.patchify8_u8(u8d, got)
    ops.patchify8(_reference_ingest(u8d).contiguous(), want)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    # and against plain torch unfold of the normalised image (K order c, kh, kw)
    ref = torch.nn.functional.unfold(_reference_ingest(u8d), kernel_size=8, stride=8).transpose(1, 2).reshape(M, 192)
    assert torch.equal(got.float(), ref.to(torch.bfloat16).float())


def test_patchify_u8_rejects_bad_input(cuda_device):
    out = torch.empty(1, 192, device=cuda_device, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.patchify8_u8(torch.zeros(1, 8, 8, 4, device=cuda_device, dtype=torch.uint8), out)
    with pytest.raises(RuntimeError):
        ops.patchify8_u8(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), out)          # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        ops.patchify8_u8(torch.zeros(1, 12, 8, 3, device=cuda_device, dtype=torch.uint8), out)   # H % 8 != 0


@pytest.mark.parametrize("cfg_name,batch", [("vit-tiny-test", 3), ("vit-s-vqgan", 2)])
def test_pixels_in_pixels_out_equals_reference_transform_chain(cuda_device, cfg_name, batch):
    cfg, sd, _ = seeded_vqgan(cfg_name, 3)
    model = _model(cfg_name, sd, cuda_device)
    S = cfg["enc"]["image_size"]
    g = torch.Generator().manual_seed(11)
    u8 = torch.randint(0, 256, (batch, S, S, 3), generator=g, dtype=torch.uint8).to(cuda_device)

    # ingest: encode_pixels(u8) == encode(Normalize(ToTensor(u8))) — identical kernels downstream, identical bits
    zq_a, loss_a, idx_a = model.encode_pixels(u8)
    zq_b, loss_b, idx_b = model.encode(_reference_ingest(u8))
    assert torch.equal(idx_a, idx_b) and torch.equal(zq_a, zq_b) and torch.equal(loss_a, loss_b)

    # egress: decode_pixels(z) == restore(decode(z)), bit-exact uint8
    px = model.decode_pixels(zq_a)
    assert px.dtype == torch.uint8 and px.shape == (batch, S, S, 3)
    want = _reference_restore(model.decode(zq_a))
    np.testing.assert_array_equal(px.cpu().numpy(), want)
    px2 = model.decode_pixels_from_indice(idx_a)
    want2 = _reference_restore(model.decode_from_indice(idx_a))
    np.testing.assert_array_equal(px2.cpu().numpy(), want2)
    # saturation: the clamp maps to 0 / 255 exactly
    assert int(px.max()) <= 255 and int(px.min()) >= 0


def test_encode_pixels_type_errors(cuda_device):
    cfg, sd, _ = seeded_vqgan("vit-tiny-test", 3)
    model = _model("vit-tiny-test", sd, cuda_device)
    S = cfg["enc"]["image_size"]
    with pytest.raises(TypeError):
        model.encode_pixels(torch.zeros(1, S, S, 3, device=cuda_device))
    with pytest.raises(RuntimeError):
        model.encode_pixels(torch.zeros(1, S, S + 8, 3, device=cuda_device, dtype=torch.uint8))


def test_unpatchify_store_modes_agree_with_einops_layout(cuda_device):
    """The two fused stores of the last projection against the reference's rearrange
    'b (h w) (p1 p2 c) -> b c (h p1) (w p2)' (layers.py:150) on a plain GEMM: fp32 NCHW (rows of W permuted to
    (c p1 p2), see PM_OUT_UNPATCH) and uint8 NHWC (reference row order)."""
    from paintmind_b200.ops import PM_OUT_F32, PM_OUT_UNPATCH, PM_OUT_UNPATCH_U8
    B, G, P, C, K = 2, 5, 8, 3, 128
    g = torch.Generator().manual_seed(9)
    a = (torch.randn(B * G * G, K, generator=g) * 0.2).to(cuda_device).bfloat16()
    w = (torch.randn(P * P * C, K, generator=g) * 0.3).to(cuda_device).bfloat16()
    bias = (torch.randn(P * P * C, generator=g) * 0.2).to(cuda_device)
    flat = torch.empty(B * G * G, P * P * C, device=cuda_device)
    ops.gemm(a, w, flat, bias=bias, out_mode=PM_OUT_F32)
    want = flat.view(B, G, G, P, P, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, G * P, G * P).clamp(-1, 1)
    perm = torch.arange(P * P * C, device=cuda_device).view(P, P, C).permute(2, 0, 1).reshape(-1)
    img = torch.full((B, C, G * P, G * P), float("nan"), device=cuda_device)
    ops.gemm(a, w[perm].contiguous(), img, bias=bias[perm].contiguous(), out_mode=PM_OUT_UNPATCH, patch=P, channels=C, grid=G)
    assert torch.equal(img, want)
    px = torch.zeros(B, G * P, G * P, C, device=cuda_device, dtype=torch.uint8)
    ops.gemm(a, w, px, bias=bias, out_mode=PM_OUT_UNPATCH_U8, patch=P, channels=C, grid=G)
    np.testing.assert_array_equal(px.cpu().numpy(), _reference_restore(want))
