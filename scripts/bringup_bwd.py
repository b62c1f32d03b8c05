"""GPU bring-up of the backward-path kernels: each kernel against torch autograd in fp32 on the
same bf16-rounded inputs, plus timings at the batch-256 shapes.  Run on the B200 box: python scripts/bringup_bwd.py [quick]
(writes gpurun_out/bringup_bwd.log)."""
import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
out_dir = Path("gpurun_out"); out_dir.mkdir(exist_ok=True)
logf = open(out_dir / "bringup_bwd.log", "w")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    logf.write(s + "\n"); logf.flush()


def report(name, got, ref, tol):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    scale = max(ref.abs().max().item(), 1e-20)
    rel = err.max().item() / scale
    nrm = (got - ref).norm().item() / max(ref.norm().item(), 1e-20)
    ok = rel <= tol and math.isfinite(rel)
    log(f"[{'OK ' if ok else 'BAD'}] {name}: max_err/absmax={rel:.3e} rel_l2={nrm:.3e} absmax={scale:.4g}")
    return ok


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def want(name):
    return only is None or name in only


g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale)


# --------------------------------------------------------------------------------------------- wgrad
if want("wgrad"):
    for (M, N, K) in [(4096, 1536, 512), (8192, 512, 1408), (2048, 64, 512), (2048, 512, 64), (2048, 192, 512), (2048, 512, 192),
                      (4096 + 64, 2816, 512), (1000, 128, 256)]:
        dy = rnd(M, N).to(torch.bfloat16); x = rnd(M, K).to(torch.bfloat16)
        out = torch.full((N, K), 7.0, device=dev)
        ops.wgrad(dy, x, out)
        ref = dy.float().t() @ x.float()
        report(f"wgrad M={M} N={N} K={K}", out, ref, 2e-3)
        ops.wgrad(dy, x, out, accumulate=True)
        report(f"wgrad accumulate M={M} N={N} K={K}", out, 2 * ref, 2e-3)
    # sub-view operands (q|k|v style column slices) and padded N
    M = 4096
    big = rnd(M, 1536).to(torch.bfloat16); x = rnd(M, 512).to(torch.bfloat16)
    out = torch.empty(512, 512, device=dev)
    ops.wgrad(big[:, 512:1024], x, out)
    report("wgrad column-slice dy", out, big[:, 512:1024].float().t() @ x.float(), 2e-3)
    if not quick:
        M = 262144
        for (N, K) in [(1536, 512), (512, 512), (2816, 512), (512, 1408), (192, 512), (512, 192)]:
            dy = rnd(M, N).to(torch.bfloat16); x = rnd(M, K).to(torch.bfloat16)
            out = torch.empty(N, K, device=dev)
            ms = timeit(lambda: ops.wgrad(dy, x, out))
            log(f"   wgrad M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")
            del dy, x

# --------------------------------------------------------------------------------------------- colsum
if want("colsum"):
    for (M, N) in [(4096, 512), (8192, 2816), (256, 1024 * 512), (1000, 64)]:
        x = rnd(M, N).to(torch.bfloat16)
        out = torch.empty(N, device=dev)
        ops.colsum(x, out)
        report(f"colsum M={M} N={N}", out, x.float().sum(0), 1e-4)
    if not quick:
        x = rnd(262144, 512).to(torch.bfloat16); out = torch.empty(512, device=dev)
        ms = timeit(lambda: ops.colsum(x, out))
        log(f"   colsum 262144x512: {ms:.3f} ms  {x.numel() * 2 / ms / 1e6:.0f} GB/s")

# --------------------------------------------------------------------------------------------- layernorm backward
if want("ln"):
    for (M, D) in [(4096, 512), (2048, 1024), (1000, 256)]:
        x = rnd(M, D).to(torch.bfloat16); dn = rnd(M, D).to(torch.bfloat16); dres = rnd(M, D).to(torch.bfloat16)
        gamma = (1 + 0.1 * rnd(D)).contiguous(); beta = 0.1 * rnd(D)
        xr = x.float().requires_grad_(True); gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
        y = F.layer_norm(xr, (D,), gr, br, 1e-5)
        y.backward(dn.float())
        dx = torch.empty(M, D, device=dev, dtype=torch.bfloat16); dgb = torch.empty(2, D, device=dev)
        ops.layernorm_bwd(dn, x, gamma, dx, dgb, dres=dres)
        report(f"ln_bwd dx M={M} D={D}", dx, xr.grad + dres.float(), 1e-2)
        report(f"ln_bwd dgamma M={M} D={D}", dgb[0], gr.grad, 1e-3)
        report(f"ln_bwd dbeta M={M} D={D}", dgb[1], br.grad, 1e-3)
    if not quick:
        M, D = 262144, 512
        x = rnd(M, D).to(torch.bfloat16); dn = rnd(M, D).to(torch.bfloat16); dres = rnd(M, D).to(torch.bfloat16)
        gamma = torch.ones(D, device=dev); dx = torch.empty_like(x); dgb = torch.empty(2, D, device=dev)
        ms = timeit(lambda: ops.layernorm_bwd(dn, x, gamma, dx, dgb, dres=dres))
        log(f"   ln_bwd 262144x512: {ms:.3f} ms  {4 * x.numel() * 2 / ms / 1e6:.0f} GB/s")

# --------------------------------------------------------------------------------------------- swiglu backward
if want("swiglu"):
    M, hp = 2048, 1408
    x12 = rnd(M, 2 * hp).to(torch.bfloat16); dh = rnd(M, hp).to(torch.bfloat16)
    h = torch.empty(M, hp, device=dev, dtype=torch.bfloat16); d12 = torch.empty(M, 2 * hp, device=dev, dtype=torch.bfloat16)
    b12 = torch.empty(2 * hp, device=dev)
    ops.swiglu_bwd(x12, dh, h, d12, b12)
    t = x12.float().view(M, hp // 128, 2, 128)
    gg = t[:, :, 0].reshape(M, hp).clone().requires_grad_(True); vv = t[:, :, 1].reshape(M, hp).clone().requires_grad_(True)
    hr = F.silu(gg) * vv
    hr.backward(dh.float())
    report("swiglu_bwd h", h, hr, 1e-2)
    dref = torch.stack([gg.grad.view(M, hp // 128, 128), vv.grad.view(M, hp // 128, 128)], dim=2).reshape(M, 2 * hp)
    report("swiglu_bwd d12", d12, dref, 1e-2)
    report("swiglu_bwd b12 (fused column sums)", b12, dref.sum(0), 1e-4)
    if not quick:
        M = 262144
        x12 = rnd(M, 2 * hp).to(torch.bfloat16); dh = rnd(M, hp).to(torch.bfloat16)
        h = torch.empty(M, hp, device=dev, dtype=torch.bfloat16); d12 = torch.empty(M, 2 * hp, device=dev, dtype=torch.bfloat16)
        b12 = torch.empty(2 * hp, device=dev)
        ms = timeit(lambda: ops.swiglu_bwd(x12, dh, h, d12, b12))
        log(f"   swiglu_bwd 262144x1408: {ms:.3f} ms  {6 * M * hp * 2 / ms / 1e6:.0f} GB/s")
        del x12, dh, h, d12

# --------------------------------------------------------------------------------------------- attention forward lse + backward
if want("attn"):
    for (B, H, N) in [(2, 8, 1024), (1, 2, 128), (3, 8, 256), (2, 2, 64), (2, 3, 200)]:
        inner = H * 64
        qkv = rnd(B, N, 3 * inner, scale=1.0).to(torch.bfloat16)
        q, k, v = qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
        o = torch.empty(B, N, inner, device=dev, dtype=torch.bfloat16)
        lse = ops.lse_buffer(B, H, N, dev)
        scale = 0.125
        ops.attention_train(q, k, v, o, H, scale, lse)
        qr, kr, vr = [t.float().view(B, N, H, 64).permute(0, 2, 1, 3).contiguous().requires_grad_(True) for t in (q, k, v)]
        sim = (qr * scale) @ kr.transpose(-1, -2)
        lse_ref = torch.logsumexp(sim, dim=-1) * 1.4426950408889634
        oref = torch.softmax(sim, dim=-1) @ vr
        report(f"attn fwd o B={B} H={H} N={N}", o.view(B, N, H, 64).permute(0, 2, 1, 3), oref, 2e-2)
        report(f"attn fwd lse B={B} H={H} N={N}", lse, lse_ref, 1e-4)
        do = rnd(B, N, inner).to(torch.bfloat16)
        oref.backward(do.float().view(B, N, H, 64).permute(0, 2, 1, 3))
        dqkv = torch.zeros(B, N, 3 * inner, device=dev, dtype=torch.bfloat16)
        if N == 256:
            o32 = torch.empty(B, N, inner, device=dev)
            ops.attention_train(q, k, v, o, H, scale, lse, o32)
            report(f"attn fwd o32 B={B} H={H} N={N}", o32.view(B, N, H, 64).permute(0, 2, 1, 3), oref, 2e-2)
        ops.attention_bwd(q, k, v, o, do, lse, dqkv[..., :inner], dqkv[..., inner:2 * inner], dqkv[..., 2 * inner:], H, scale)
        torch.cuda.synchronize()
        for nm, got, ref in (("dq", dqkv[..., :inner], qr.grad), ("dk", dqkv[..., inner:2 * inner], kr.grad), ("dv", dqkv[..., 2 * inner:], vr.grad)):
            report(f"attn bwd {nm} B={B} H={H} N={N}", got.reshape(B, N, H, 64).permute(0, 2, 1, 3), ref, 3e-2)
    if not quick:
        B, H, N = 256, 8, 1024
        inner = 512
        qkv = rnd(B, N, 3 * inner).to(torch.bfloat16)
        q, k, v = qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
        o = torch.empty(B, N, inner, device=dev, dtype=torch.bfloat16); lse = ops.lse_buffer(B, H, N, dev)
        do = rnd(B, N, inner).to(torch.bfloat16); dqkv = torch.empty_like(qkv)
        ms = timeit(lambda: ops.attention_train(q, k, v, o, H, 0.125, lse))
        log(f"   attn fwd (+lse) B=256: {ms:.3f} ms  {4.0 * B * H * N * N * 64 / ms / 1e9:.0f} TFLOP/s")
        ms = timeit(lambda: ops.attention_bwd(q, k, v, o, do, lse, dqkv[..., :inner], dqkv[..., inner:2 * inner], dqkv[..., 2 * inner:], H, 0.125))
        log(f"   attn bwd B=256: {ms:.3f} ms  {14.0 * B * H * N * N * 64 / ms / 1e9:.0f} TFLOP/s (7 matmuls incl. recompute)")

# --------------------------------------------------------------------------------------------- vq backward
if want("vq"):
    M, n_e = 4096, 8192
    z = rnd(M, 32); E = rnd(n_e, 32); d_out = rnd(M, 32, scale=0.01); d_loss = torch.tensor(0.7, device=dev)
    beta = 0.25
    zr = z.clone().requires_grad_(True); Er = E.clone().requires_grad_(True)
    zn = F.normalize(zr, dim=-1); en_all = F.normalize(Er, dim=-1)
    idx = torch.argmax(zn.detach() @ en_all.detach().t(), dim=1)
    zq = F.normalize(Er[idx], dim=-1)
    loss = beta * torch.mean((zq.detach() - zn) ** 2) + torch.mean((zq - zn.detach()) ** 2)
    outv = zn + (zq - zn).detach()
    (outv * d_out).sum().add(loss * d_loss).backward()
    dz = torch.empty(M, 32, device=dev); dzs = torch.empty(M, 64, device=dev, dtype=torch.bfloat16); dE = torch.zeros(n_e, 32, device=dev)
    ops.vq_bwd(z, idx, E, d_out, d_loss, beta, dz=dz, dz_split=dzs, dE=dE)
    report("vq_bwd dz", dz, zr.grad, 1e-4)
    report("vq_bwd dz_split", dzs[:, :32].float() + dzs[:, 32:].float(), zr.grad, 1e-4)
    report("vq_bwd dE", dE, Er.grad, 1e-4)

# --------------------------------------------------------------------------------------------- unpatchify backward
if want("unpatch"):
    B = 4
    gimg = rnd(B, 3, 256, 256); rec = rnd(B, 3, 256, 256).clamp(-1, 1)
    out = torch.empty(B * 1024, 192, device=dev, dtype=torch.bfloat16)
    ops.unpatchify8_bwd(gimg, rec, out)
    m = (rec.abs() < 1).float() * gimg
    ref = m.view(B, 3, 32, 8, 32, 8).permute(0, 2, 4, 1, 3, 5).reshape(B * 1024, 192)
    report("unpatchify8_bwd", out, ref, 1e-2)
log("done")
