"""The pre-scaled-query attention kernels (pm_attn4.cu, 16 softmax warps, the default; pm_attn3.cu, 8 warps: row maximum
subtracted by the tensor core, no per-tile maximum after the first key tile) against fp32 softmax attention (reference modules/attention.py:51-58), including the inputs its
short-cuts must survive: ragged key / query counts (mask through the bias operand), logits tens of binades above the first key tile's maximum (stale-maximum path) and by more than fp32's exponent range (overflow -> exact re-run of the work item)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

LOG2E = 1.4426950408889634


def _ref(qp, k, v, heads):
    """softmax over base-2 logits qp k^T (qp carries scale * log2 e), fp64 on the bf16 operands."""
    B, Nq, _ = qp.shape
    Nk = k.shape[1]
    qf = qp.double().view(B, Nq, heads, 64).transpose(1, 2)
    kf = k.double().view(B, Nk, heads, 64).transpose(1, 2)
    vf = v.double().view(B, Nk, heads, 64).transpose(1, 2)
    s = qf @ kf.transpose(-1, -2) * math.log(2.0)
    return (torch.softmax(s, dim=-1) @ vf).transpose(1, 2).reshape(B, Nq, heads * 64).float()


def _run(qp, k, v, heads):
    from paintmind_b200 import ops
    o = torch.full((qp.shape[0], qp.shape[1], heads * 64), float("nan"), device=qp.device, dtype=torch.bfloat16)
    ops.attention(qp, k, v, o, heads, 1.0, prescaled=True)
    torch.cuda.synchronize()
    return o.float()


@pytest.mark.parametrize("B,Nq,Nk,heads", [(3, 1024, 1024, 8), (2, 1024, 77, 16), (3, 200, 200, 2), (1, 100, 64, 2), (5, 256, 640, 4),
                                           (160, 256, 256, 1)])
def test_prescaled_attention_vs_fp32_softmax(cuda_device, B, Nq, Nk, heads):
    torch.manual_seed(B * 1000 + Nk)
    inner = heads * 64
    q = torch.randn(B, Nq, inner, device=cuda_device)
    k = torch.randn(B, Nk, inner, device=cuda_device).bfloat16()
    v = torch.randn(B, Nk, inner, device=cuda_device).bfloat16()
    qp = (q * (0.125 * LOG2E)).bfloat16()
    got = _run(qp, k, v, heads)
    ref = _ref(qp, k, v, heads)
    err = (got - ref).abs()
    assert torch.isfinite(got).all()
    assert err.max().item() < 0.02 and err.mean().item() < 0.002, (err.max().item(), err.mean().item())


def test_prescaled_attention_views_of_a_packed_qkv_buffer(cuda_device):
    """q | k | v read in place from one token-major buffer (how engine.py calls it), output into a wider buffer."""
    from paintmind_b200 import ops
    torch.manual_seed(3)
    B, N, H = 2, 384, 4
    qkv = torch.randn(B, N, 3 * 256, device=cuda_device)
    qkv[..., :256] *= 0.125 * LOG2E
    qkv = qkv.bfloat16()
    o = torch.zeros(B, N, 256, device=cuda_device, dtype=torch.bfloat16)
    ops.attention(qkv[..., :256], qkv[..., 256:512], qkv[..., 512:], o, H, 1.0, prescaled=True)
    ref = _ref(qkv[..., :256], qkv[..., 256:512], qkv[..., 512:], H)
    assert (o.float() - ref).abs().max().item() < 0.02


@pytest.mark.parametrize("gain", [2.0, 3.0])
def test_prescaled_attention_large_logits_recentre(cuda_device, gain):
    """Base-2 logits with a standard deviation of 46-100: later key tiles exceed the first tile's maximum by tens of binades
    (P far above 1, rows beyond 2^100 are re-run exactly) — results must still match."""
    torch.manual_seed(11)
    B, N, H = 3, 640, 2
    q = (torch.randn(B, N, 128, device=cuda_device) * gain * LOG2E).bfloat16()
    k = (torch.randn(B, N, 128, device=cuda_device) * gain).bfloat16()
    v = torch.randn(B, N, 128, device=cuda_device).bfloat16()
    got = _run(q, k, v, H)
    ref = _ref(q, k, v, H)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 0.03


def test_prescaled_attention_overflow_falls_back_to_exact_pass(cuda_device):
    """One key far down the sequence whose logit is > 2^127 above everything in the first tile: the fast path overflows (inf
    row sum), the item is flagged and re-run with per-tile maxima.  Rows of other items are untouched."""
    torch.manual_seed(5)
    B, N, H = 2, 1024, 2
    q = (torch.randn(B, N, 128, device=cuda_device) * 0.2).bfloat16()
    k = (torch.randn(B, N, 128, device=cuda_device) * 0.2).bfloat16()
    v = torch.randn(B, N, 128, device=cuda_device).bfloat16()
    # batch 1, head 0: queries 300..399 are strongly aligned with key 700 (logit ~ +1000 in base 2)
    q[1, 300:400, :64] = 4.0
    k[1, 700, :64] = 4.0
    got = _run(q, k, v, H)
    ref = _ref(q, k, v, H)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 0.02
    assert (got[1, 300:400, :64] - v[1, 700, :64].float()[None]).abs().max().item() < 1e-2


def test_prescaled_matches_unscaled_kernel(cuda_device):
    """Same inputs through the round-1 kernel (scale applied per score) and the pre-scaled kernel: both within bf16 noise."""
    from paintmind_b200 import ops
    torch.manual_seed(7)
    B, N, H = 4, 512, 8
    q = torch.randn(B, N, 512, device=cuda_device)
    k = torch.randn(B, N, 512, device=cuda_device).bfloat16()
    v = torch.randn(B, N, 512, device=cuda_device).bfloat16()
    o1 = torch.empty(B, N, 512, device=cuda_device, dtype=torch.bfloat16)
    ops.attention(q.bfloat16(), k, v, o1, H, 0.125)
    o3 = _run((q * (0.125 * LOG2E)).bfloat16(), k, v, H)
    assert (o1.float() - o3).abs().max().item() < 0.02


def test_eight_warp_variant_in_a_subprocess(cuda_device):
    """PM_ATTN_PRE=3 selects pm_attn3.cu (8 softmax warps; the library reads the variable once per process): same checks, one
    child process — self-attention, ragged 77-key cross-attention, several items per CTA, and the overflow re-run."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    child = r"""
import sys, math, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import test_gpu_attention3 as T
dev = torch.device("cuda:0")
for (B, Nq, Nk, heads) in [(3, 1024, 1024, 8), (2, 1024, 77, 16), (160, 256, 256, 1), (3, 200, 200, 2)]:
    T.test_prescaled_attention_vs_fp32_softmax(dev, B, Nq, Nk, heads)
T.test_prescaled_attention_large_logits_recentre(dev, 2.0)
T.test_prescaled_attention_overflow_falls_back_to_exact_pass(dev)
print("CHILD OK")
""" % (str(root), str(root / "tests"))
    r = subprocess.run([sys.executable, "-c", child], env=dict(os.environ, PM_ATTN_PRE="3"), capture_output=True, text=True, timeout=300)
    assert "CHILD OK" in r.stdout, r.stdout[-800:] + r.stderr[-1500:]


def test_prescaled_attention_is_batch_invariant(cuda_device):
    """A row's output bits do not depend on which other work items the CTA processed before it (every item's first key tile is
    issued with offset 0): the same samples inside a larger batch — several items per CTA, another item -> CTA assignment —
    give bit-identical outputs."""
    torch.manual_seed(13)
    B, N, H = 40, 512, 8
    q = (torch.randn(B, N, 512, device=cuda_device) * (0.125 * LOG2E)).bfloat16()
    k = torch.randn(B, N, 512, device=cuda_device).bfloat16()
    v = torch.randn(B, N, 512, device=cuda_device).bfloat16()
    full = _run(q, k, v, H)
    part = _run(q[3:9].contiguous(), k[3:9].contiguous(), v[3:9].contiguous(), H)
    assert torch.equal(full[3:9], part)
