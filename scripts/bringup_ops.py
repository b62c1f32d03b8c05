"""GPU bring-up of attention / VQ / row kernels vs torch references."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
out_dir = Path("gpurun_out"); out_dir.mkdir(exist_ok=True)
logf = open(out_dir / "bringup_ops.log", "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); logf.write(s + "\n"); logf.flush()


def report(name, got, ref, tol):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    mx = err.max().item(); scale = max(ref.abs().max().item(), 1e-6)
    ok = bool(mx <= tol * max(scale, 1.0)) and bool(torch.isfinite(got).all())
    log(f"[{'OK ' if ok else 'BAD'}] {name}: max_err={mx:.4g} ref_absmax={scale:.4g} mean_err={err.mean().item():.4g}")
    if not ok:
        N = err.shape[-1]
        e2 = err.reshape(-1, N); bad = e2 > tol * max(scale, 1.0)
        log(f"   bad frac={bad.float().mean().item():.4f} rows={bad.any(1).nonzero().flatten()[:12].tolist()} cols={bad.any(0).nonzero().flatten()[:12].tolist()}")
        log("   got[0,:8]=", got.reshape(-1, N)[0, :8].tolist()); log("   ref[0,:8]=", ref.reshape(-1, N)[0, :8].tolist())
        log("   nan count", torch.isnan(got).sum().item())
    return ok


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def attn_ref(q, k, v, H, scale):
    B, Nq, _ = q.shape; Nk = k.shape[1]
    qh = q.float().reshape(B, Nq, H, 64).transpose(1, 2); kh = k.float().reshape(B, Nk, H, 64).transpose(1, 2)
    vh = v.float().reshape(B, Nk, H, 64).transpose(1, 2)
    s = (qh * scale) @ kh.transpose(-1, -2)
    return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Nq, H * 64)


def main():
    ok = True
    torch.manual_seed(0)
    # ---------------- attention ----------------
    for (B, H, Nq, Nk, sc) in [(1, 1, 128, 128, 1.0), (2, 2, 256, 384, 1.0), (2, 8, 1024, 1024, 1.0), (2, 4, 1024, 77, 2.0),
                               (1, 2, 64, 64, 3.0), (3, 2, 200, 130, 1.0)]:
        inner = H * 64
        qkv = (torch.randn(B, Nq, 3 * inner, device=dev) * sc).bfloat16()
        kv = (torch.randn(B, Nk, 2 * inner, device=dev) * sc).bfloat16()
        q = qkv[..., :inner]
        if Nk == Nq:
            k, v = qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
        else:
            k, v = kv[..., :inner], kv[..., inner:]
        o = torch.full((B, Nq, inner), float("nan"), device=dev, dtype=torch.bfloat16)
        try:
            ops.attention(q, k, v, o, H, 0.125)
            torch.cuda.synchronize()
            ok &= report(f"attn B{B} H{H} Nq{Nq} Nk{Nk} sc{sc}", o, attn_ref(q, k, v, H, 0.125), 2e-2)
        except Exception as e:  # noqa: BLE001
            log("attn EXC", repr(e)); ok = False; break
    # perf at B=256 (one layer): 8 heads x 1024 tokens
    try:
        B, H, N = 256, 8, 1024
        qkv = torch.randn(B, N, 3 * 512, device=dev).bfloat16()
        o = torch.empty(B, N, 512, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.attention(qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:], o, H, 0.125))
        fl = 4.0 * B * H * N * N * 64
        log(f"[perf] attention B256 H8 N1024: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
        ms2 = timeit(lambda: F.scaled_dot_product_attention(qkv[..., :512].reshape(B, N, H, 64).transpose(1, 2),
                                                            qkv[..., 512:1024].reshape(B, N, H, 64).transpose(1, 2),
                                                            qkv[..., 1024:].reshape(B, N, H, 64).transpose(1, 2)))
        log(f"        torch SDPA same shape: {ms2:.3f} ms")
    except Exception as e:  # noqa: BLE001
        log("attn perf EXC", repr(e)); ok = False

    # ---------------- layernorm / stats / patchify ----------------
    for (M, D) in [(1000, 512), (64, 128), (513, 1024)]:
        x = (torch.randn(M, D, device=dev) * 2 + 0.3).bfloat16()
        g = torch.rand(D, device=dev) + 0.5; b = torch.randn(D, device=dev) * 0.1
        stats = torch.empty(M, 2, device=dev)
        ops.layernorm(x, stats=stats)
        xf = x.float(); mu = xf.mean(1); rstd = (xf.var(1, unbiased=False) + 1e-5).rsqrt()
        ok &= report(f"ln_stats M{M} D{D}", stats, torch.stack([mu, rstd], 1), 1e-4)
        y = torch.empty_like(x); st2 = torch.empty(M, 2, device=dev)
        ops.layernorm(x, gamma=g, beta=b, y=y, stats=st2)
        ok &= report(f"layernorm M{M} D{D}", y, F.layer_norm(xf, (D,), g, b, 1e-5), 1e-2)
        yf = y.float()
        ok &= report("   stats of y", st2, torch.stack([yf.mean(1), (yf.var(1, unbiased=False) + 1e-5).rsqrt()], 1), 1e-4)
    img = torch.rand(3, 3, 64, 96, device=dev) * 2 - 1
    outp = torch.empty(3 * 8 * 12, 192, device=dev, dtype=torch.bfloat16)
    ops.patchify8(img, outp)
    refp = img.reshape(3, 3, 8, 8, 12, 8).permute(0, 2, 4, 1, 3, 5).reshape(-1, 192)
    ok &= report("patchify8", outp, refp.bfloat16(), 1e-6)

    # ---------------- VQ ----------------
    for (M, n_e, splits) in [(256, 512, 1), (1000, 8192, 1), (1000, 8192, 4), (4096, 8192, 0), (65536, 8192, 0), (65536, 8192, 1), (300, 1000, 1)]:
        g = torch.Generator(device="cpu").manual_seed(0)
        z = torch.randn(M, 32, generator=g).to(dev) * 3.0
        E = torch.randn(n_e, 32, generator=g).to(dev)
        try:
            en, packed = ops.vq_codebook_prep(E)
            idx = torch.full((M,), -1, device=dev, dtype=torch.int64)
            zq = torch.empty(M, 32, device=dev); zs = torch.empty(M, 64, device=dev, dtype=torch.bfloat16)
            sse = torch.zeros(1, device=dev, dtype=torch.float64); hist = torch.zeros(n_e, device=dev, dtype=torch.int64)
            cv = torch.empty(8, M, device=dev); ci = torch.empty(8, M, device=dev, dtype=torch.int32)
            ops.vq_forward(z, en, packed, idx=idx, zq=zq, zq_split=zs, sse=sse, hist=hist, cand_val=cv, cand_idx=ci, splits=splits)
            torch.cuda.synchronize()
            zn = F.normalize(z, dim=-1); enr = F.normalize(E, dim=-1)
            ok &= report(f"   en M{M} n_e{n_e}", en, enr, 1e-6)
            d = (zn ** 2).sum(1, keepdim=True) + (enr ** 2).sum(1) - 2 * zn @ enr.t()
            ref_idx = d.argmin(1)
            top2 = d.topk(2, dim=1, largest=False).values; gap = top2[:, 1] - top2[:, 0]
            mism = (idx != ref_idx)
            nm = int(mism.sum().item())
            worst = gap[mism].max().item() if nm else 0.0
            # exact check in float64
            d64 = -(zn.double() @ enr.double().t()); ref64 = d64.argmin(1)
            nm64 = int((idx != ref64).sum().item())
            good = nm == 0 or worst < 1e-5
            ok &= good
            log(f"[{'OK ' if good else 'BAD'}] vq idx M{M} n_e{n_e} splits{splits}: mismatches vs fp32 argmin={nm} (max gap at mismatch {worst:.3g}), vs fp64 argmax={nm64}, min gap overall={gap.min().item():.3g}")
            zq_ref = zn + (enr[ref_idx] - zn)
            sel = ~mism
            ok &= report("   zq", zq[sel], zq_ref[sel], 1e-6)
            ok &= report("   zq_split hi+lo", zs[:, :32].float() + zs[:, 32:].float(), zq, 1e-5)
            sse_ref = ((enr[idx] - zn).double() ** 2).sum().item()
            rel = abs(sse.item() - sse_ref) / max(sse_ref, 1e-30)
            log(f"   sse rel err {rel:.3g}; hist ok {bool((hist == torch.bincount(idx, minlength=n_e)).all())}")
            ok &= rel < 1e-5 and bool((hist == torch.bincount(idx, minlength=n_e)).all())
        except Exception as e:  # noqa: BLE001
            log("vq EXC", repr(e)); ok = False; break
    try:
        M, n_e = 65536, 8192
        g = torch.Generator(device="cpu").manual_seed(0)
        z = F.normalize(torch.randn(M, 32, generator=g), dim=-1).to(dev); E = torch.randn(n_e, 32, generator=g).to(dev)
        en, packed = ops.vq_codebook_prep(E)
        idx = torch.empty(M, device=dev, dtype=torch.int64); zq = torch.empty(M, 32, device=dev)
        sse = torch.zeros(1, device=dev, dtype=torch.float64)
        cv = torch.empty(8, M, device=dev); ci = torch.empty(8, M, device=dev, dtype=torch.int32)
        for sp in (0, 1, 2, 4, 8):
            ms = timeit(lambda: ops.vq_forward(z, en, packed, idx=idx, zq=zq, sse=sse, cand_val=cv, cand_idx=ci, splits=sp), iters=20)
            log(f"[perf] vq 65536x8192 splits={sp}: {ms * 1e3:.1f} us  {M / ms / 1e6:.2f} G lookups/s  ({2.0 * M * n_e * 32 / ms / 1e9:.1f} nominal TFLOP/s)")
        M2 = 262144
        z2 = torch.randn(M2, 32, device=dev); idx2 = torch.empty(M2, device=dev, dtype=torch.int64); zq2 = torch.empty(M2, 32, device=dev)
        ms = timeit(lambda: ops.vq_forward(z2, en, packed, idx=idx2, zq=zq2, sse=sse, splits=1), iters=10)
        log(f"[perf] vq 262144x8192 fused: {ms:.3f} ms  {M2 / ms / 1e6:.2f} G lookups/s")
    except Exception as e:  # noqa: BLE001
        log("vq perf EXC", repr(e)); ok = False

    # gather / split
    idxg = torch.randint(0, 8193, (500,), device=dev)
    table = torch.randn(8193, 32, device=dev)
    o1 = torch.empty(500, 32, device=dev); o2 = torch.empty(500, 64, device=dev, dtype=torch.bfloat16)
    ops.vq_gather(idxg, table, True, o1, o2)
    ok &= report("gather+l2norm", o1, F.normalize(table[idxg], dim=-1), 1e-6)
    ops.vq_gather(idxg, table, False, o1, None)
    ok &= report("gather raw", o1, table[idxg], 0)
    ops.split_rows32(table[:500], o2)
    ok &= report("split_rows32", o2[:, :32].float() + o2[:, 32:].float(), table[:500], 1e-5)
    log("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    try:
        rc = main()
    except Exception as e:  # noqa: BLE001
        import traceback
        log("EXCEPTION:", repr(e)); log(traceback.format_exc())
        rc = 2
    sys.exit(rc)
