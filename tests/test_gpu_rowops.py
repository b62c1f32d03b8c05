"""GPU parity of the LayerNorm row kernels at op level (through the C-ABI): every kernel variant (register-resident fast
path for D = 256 / 512 / 1024, generic path for other widths), ragged row counts (odd M, M smaller than a block, M larger
than one wave of the persistent grid), stats-only mode, and the backward kernel against torch autograd.

Reference semantics: nn.LayerNorm (stage1/layers.py:49,51,89,128), eps inside the square root, biased variance.
Tolerances: the forward output is bf16 — at most one bf16 ulp (2^-8 relative) from the fp32 result rounded once; statistics
and gradients fp32 accumulate over D <= 1024 terms (1e-5 / 2e-2 relative L2, the latter bounded by the bf16 inputs)."""
import pytest
import torch
import torch.nn.functional as F

from paintmind_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D", [64, 256, 512, 768, 1024])
@pytest.mark.parametrize("M", [1, 5, 33, 5000])
def test_layernorm_forward_and_stats(cuda_device, M, D):
    g = torch.Generator(device=cuda_device).manual_seed(M * 131 + D)
    x = (torch.randn(M, D, device=cuda_device, generator=g) * 1.7 + 0.3).bfloat16()
    gamma = 1 + 0.2 * torch.randn(D, device=cuda_device, generator=g)
    beta = 0.1 * torch.randn(D, device=cuda_device, generator=g)
    y = torch.full((M, D), float("nan"), device=cuda_device, dtype=torch.bfloat16)
    st = torch.full((M, 2), float("nan"), device=cuda_device)
    ops.layernorm(x, gamma=gamma, beta=beta, y=y, stats=st)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (D,), gamma, beta, 1e-5)
    err = (y.float() - ref).abs()
    assert torch.isfinite(y.float()).all()
    assert (err <= ref.abs() * 2.0 ** -8 + 1e-6).all(), float(err.max())
    # the second output: mean / rstd of the ROUNDED output row (what the LN-folded epilogue of the next GEMM consumes)
    yf = y.float()
    mean = yf.mean(dim=1)
    rstd = torch.rsqrt(yf.var(dim=1, unbiased=False) + 1e-5)
    assert torch.allclose(st[:, 0], mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(st[:, 1], rstd, rtol=1e-5, atol=1e-6)
    # stats-only mode: mean / rstd of the input row
    st0 = torch.full((M, 2), float("nan"), device=cuda_device)
    ops.layernorm(x, stats=st0)
    torch.cuda.synchronize()
    xf = x.float()
    assert torch.allclose(st0[:, 0], xf.mean(dim=1), rtol=1e-5, atol=1e-6)
    assert torch.allclose(st0[:, 1], torch.rsqrt(xf.var(dim=1, unbiased=False) + 1e-5), rtol=1e-5, atol=1e-6)


def test_layernorm_row_pitch(cuda_device):
    """x and y as column slices of wider buffers (ldx != D): the q|k|v / h operands of the engine are laid out that way."""
    M, D = 77, 512
    g = torch.Generator(device=cuda_device).manual_seed(3)
    xw = torch.randn(M, 3 * D, device=cuda_device, generator=g).bfloat16()
    yw = torch.zeros(M, 2 * D, device=cuda_device, dtype=torch.bfloat16)
    gamma = torch.rand(D, device=cuda_device, generator=g) + 0.5
    beta = torch.rand(D, device=cuda_device, generator=g) - 0.5
    x, y = xw[:, D:2 * D], yw[:, D:]
    ops.layernorm(x, gamma=gamma, beta=beta, y=y)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (D,), gamma, beta, 1e-5)
    assert ((y.float() - ref).abs() <= ref.abs() * 2.0 ** -8 + 1e-6).all()
    assert (yw[:, :D] == 0).all()


@pytest.mark.parametrize("M,D", [(1, 512), (7, 512), (1001, 512), (40, 64), (333, 256), (129, 1024), (3000, 512)])
@pytest.mark.parametrize("with_res", [False, True])
def test_layernorm_backward_vs_autograd(cuda_device, M, D, with_res):
    g = torch.Generator(device=cuda_device).manual_seed(M + D)
    x = torch.randn(M, D, device=cuda_device, generator=g).bfloat16()
    dn = torch.randn(M, D, device=cuda_device, generator=g).bfloat16()
    dres = torch.randn(M, D, device=cuda_device, generator=g).bfloat16() if with_res else None
    gamma = (1 + 0.1 * torch.randn(D, device=cuda_device, generator=g)).contiguous()
    dx = torch.full((M, D), float("nan"), device=cuda_device, dtype=torch.bfloat16)
    dgb = torch.full((2, D), float("nan"), device=cuda_device)
    ops.layernorm_bwd(dn, x, gamma, dx, dgb, dres=dres)
    torch.cuda.synchronize()
    xr = x.float().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = torch.zeros(D, device=cuda_device, requires_grad=True)
    F.layer_norm(xr, (D,), gr, br, 1e-5).backward(dn.float())
    want_dx = xr.grad + (dres.float() if with_res else 0)

    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-30))
    assert rel(dx.float(), want_dx) < 2e-2            # dx is stored in bf16
    assert rel(dgb[0], gr.grad) < 1e-4 and rel(dgb[1], br.grad) < 1e-4
