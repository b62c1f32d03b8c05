"""GPU bring-up of pm_gemm_bf16: correctness vs torch fp32 on bf16-rounded inputs + timing.

Run on the B200 box:  python scripts/bringup_gemm.py  (writes gpurun_out/bringup_gemm.log)
"""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
out_dir = Path("gpurun_out"); out_dir.mkdir(exist_ok=True)
logf = open(out_dir / "bringup_gemm.log", "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    logf.write(s + "\n"); logf.flush()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def run_gemm(a, w, out, bias=None, colsum=None, stats=None, pos=None, res=None, out_mode=0, swiglu=0, bn=0,
             patch=0, channels=0, grid=0, max_ctas=0, N=None, cta_group=1, stats_out=None, stats_raw=0):
    args = _lib.GemmArgs()
    args.a, args.w, args.out = ptr(a), ptr(w), ptr(out)
    args.bias, args.colsum, args.stats, args.pos, args.res = ptr(bias), ptr(colsum), ptr(stats), ptr(pos), ptr(res)
    args.lda, args.ldw = a.stride(0), w.stride(0)
    args.ld_out = out.stride(0) if out.dim() == 2 else 0
    args.ld_pos = pos.stride(0) if pos is not None else 0
    args.ld_res = res.stride(0) if res is not None else 0
    args.M, args.N, args.K = a.shape[0], (N if N is not None else w.shape[0]), a.shape[1]
    args.pos_rows = pos.shape[0] if pos is not None else 0
    args.out_mode, args.swiglu, args.bn = out_mode, swiglu, bn
    args.patch, args.channels, args.grid, args.max_ctas = patch, channels, grid, max_ctas
    args.cta_group, args.stats_out, args.stats_raw, args.ln_eps = cta_group, ptr(stats_out), stats_raw, 1e-5
    rc = lib.pm_gemm_bf16(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "pm_gemm_bf16")


def report(name, got, ref, tol):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    mx = err.max().item()
    scale = ref.abs().max().item()
    ok = mx <= tol * max(scale, 1.0)
    log(f"[{'OK ' if ok else 'BAD'}] {name}: max_err={mx:.4g} ref_absmax={scale:.4g} mean_err={err.mean().item():.4g}")
    if not ok:
        M, N = err.shape[-2], err.shape[-1]
        e2 = err.reshape(-1, N)
        bad = (e2 > tol * max(scale, 1.0))
        log(f"   bad fraction={bad.float().mean().item():.4f}; bad rows (first 16)={bad.any(1).nonzero().flatten()[:16].tolist()}"
            f" bad cols (first 16)={bad.any(0).nonzero().flatten()[:16].tolist()}")
        log("   got[0,:8]=", got.reshape(-1, N)[0, :8].tolist())
        log("   ref[0,:8]=", ref.reshape(-1, N)[0, :8].tolist())
        log("   got[1,:8]=", got.reshape(-1, N)[1, :8].tolist())
        log("   ref[1,:8]=", ref.reshape(-1, N)[1, :8].tolist())
        # block-wise error map (rows/8 x cols/8) for the first 64x64
        blk = e2[:64, :64].reshape(8, 8, 8, 8).amax(dim=(1, 3))
        log("   8x8-block max err (first 64x64):\n" + "\n".join("    " + " ".join(f"{x:9.3g}" for x in r) for r in blk.tolist()))
    return ok


def main():
    log("device:", torch.cuda.get_device_name(0), "check rc =", lib.pm_device_check())
    torch.manual_seed(0)
    all_ok = True

    def mk(M, N, K, scale=1.0):
        a = (torch.randn(M, K, device=dev) * scale).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        return a, w

    # 1. minimal single tile, fp32 direct store
    for (M, N, K, bn) in [(128, 64, 64, 64), (128, 32, 64, 32), (128, 64, 256, 64), (128, 128, 128, 128),
                          (256, 256, 512, 256), (384, 512, 192, 256), (200, 96, 64, 32)]:
        a, w = mk(M, N, K)
        out = torch.full((M, N), float("nan"), device=dev)
        run_gemm(a, w, out, out_mode=1, bn=bn)
        torch.cuda.synchronize()
        all_ok &= report(f"f32 M{M} N{N} K{K} bn{bn}", out, a.float() @ w.float().t(), 2e-3)

    # 2. bf16 TMA store, no residual
    for (M, N, K, bn, mc) in [(128, 64, 64, 64, 0), (256, 256, 512, 256, 0), (1024, 512, 512, 256, 3), (200, 192, 128, 64, 0),
                              (4096, 1536, 512, 256, 0)]:
        a, w = mk(M, N, K)
        out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        run_gemm(a, w, out, out_mode=0, bn=bn, max_ctas=mc)
        torch.cuda.synchronize()
        all_ok &= report(f"bf16 M{M} N{N} K{K} bn{bn} maxctas{mc}", out, a.float() @ w.float().t(), 1e-2)

    # 3. bias + residual + pos
    for (M, N, K, bn, mc) in [(128, 64, 64, 64, 0), (1024, 512, 512, 256, 2), (2048, 512, 1408, 128, 5), (4096, 512, 512, 256, 0)]:
        a, w = mk(M, N, K)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev).bfloat16()
        out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        run_gemm(a, w, out, bias=bias, res=res, out_mode=0, bn=bn, max_ctas=mc)
        torch.cuda.synchronize()
        all_ok &= report(f"bf16+bias+res M{M} N{N} K{K} bn{bn} maxctas{mc}", out, a.float() @ w.float().t() + bias + res.float(), 1e-2)
        # in-place residual (out aliases res)
        x = res.clone()
        run_gemm(a, w, x, bias=bias, res=x, out_mode=0, bn=bn, max_ctas=mc)
        torch.cuda.synchronize()
        all_ok &= report(f"   in-place residual", x, a.float() @ w.float().t() + bias + res.float(), 1e-2)
    M, N, K = 2048, 512, 192
    a, w = mk(M, N, K)
    pos = torch.randn(1024, N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    run_gemm(a, w, out, pos=pos, out_mode=0)
    torch.cuda.synchronize()
    all_ok &= report("bf16+pos", out, a.float() @ w.float().t() + pos.repeat(2, 1), 1e-2)

    # 4. LN fold
    M, N, K = 1024, 1536, 512
    x = (torch.randn(M, K, device=dev) * 2 + 0.5).bfloat16()
    gamma = torch.rand(K, device=dev) + 0.5; beta = torch.randn(K, device=dev) * 0.1
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev) * 0.1
    Wf = (W * gamma).bfloat16()
    colsum = Wf.float().sum(1)
    biasf = b + W @ beta
    xf = x.float()
    mu = xf.mean(1); var = xf.var(1, unbiased=False); rstd = (var + 1e-5).rsqrt()
    stats = torch.stack([mu, rstd], 1).contiguous()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    run_gemm(x, Wf, out, bias=biasf, colsum=colsum, stats=stats, out_mode=0)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(xf, (K,), gamma, beta, 1e-5) @ W.t() + b
    all_ok &= report("LN-fold QKV", out, ref, 2e-2)

    # 5. SwiGLU (N = 2*hid_pad, tile = 128 gate rows + 128 value rows)
    M, K, hid = 1024, 512, 1408
    a, _ = mk(M, 8, K)
    Wg = (torch.randn(hid, K, device=dev) / K ** 0.5); Wv = (torch.randn(hid, K, device=dev) / K ** 0.5)
    bg = torch.randn(hid, device=dev) * 0.1; bv = torch.randn(hid, device=dev) * 0.1
    Wp = torch.empty(2 * hid, K, device=dev); bp = torch.empty(2 * hid, device=dev)
    for t in range(hid // 128):
        Wp[t * 256:t * 256 + 128] = Wg[t * 128:(t + 1) * 128]; Wp[t * 256 + 128:(t + 1) * 256] = Wv[t * 128:(t + 1) * 128]
        bp[t * 256:t * 256 + 128] = bg[t * 128:(t + 1) * 128]; bp[t * 256 + 128:(t + 1) * 256] = bv[t * 128:(t + 1) * 128]
    Wp = Wp.bfloat16()
    out = torch.empty(M, hid, device=dev, dtype=torch.bfloat16)
    run_gemm(a, Wp, out, bias=bp, out_mode=0, swiglu=1)
    torch.cuda.synchronize()
    g = a.float() @ Wg.bfloat16().float().t() + bg; v = a.float() @ Wv.bfloat16().float().t() + bv
    all_ok &= report("SwiGLU", out, torch.nn.functional.silu(g) * v, 2e-2)

    # 6. unpatchify + clamp
    B, G, P, Cc = 2, 32, 8, 3
    M, N, K = B * G * G, P * P * Cc, 512
    a, w = mk(M, N, K)
    bias = torch.randn(N, device=dev) * 0.1
    img = torch.full((B, Cc, G * P, G * P), float("nan"), device=dev)
    perm = torch.arange(N, device=dev).view(P, P, Cc).permute(2, 0, 1).reshape(-1)      # (p1 p2 c) -> (c p1 p2) rows
    run_gemm(a, w[perm].contiguous(), img, bias=bias[perm].contiguous(), out_mode=2, patch=P, channels=Cc, grid=G)
    torch.cuda.synchronize()
    y = (a.float() @ w.float().t() + bias).reshape(B, G, G, P, P, Cc).permute(0, 5, 1, 3, 2, 4).reshape(B, Cc, G * P, G * P).clamp(-1, 1)
    all_ok &= report("unpatchify", img.reshape(-1, G * P), y.reshape(-1, G * P), 2e-3)

    # 6b. CTA pairs (tcgen05.mma.cta_group::2, 256x256 tiles)
    for (M, N, K, mc) in [(256, 256, 64, 0), (256, 256, 512, 0), (512, 512, 512, 1), (4096, 1536, 512, 0), (1000, 512, 1408, 3)]:
        a, w = mk(M, N, K)
        out = torch.full((M, N), float("nan"), device=dev)
        run_gemm(a, w, out, out_mode=1, bn=256, cta_group=2, max_ctas=mc)
        torch.cuda.synchronize()
        all_ok &= report(f"PAIR f32 M{M} N{N} K{K} maxpairs{mc}", out, a.float() @ w.float().t(), 2e-3)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev).bfloat16()
        x = res.clone()
        P = 2 * ((N + 255) // 256)
        st = torch.full((M, P, 2), float("nan"), device=dev)
        run_gemm(a, w, x, bias=bias, res=x, out_mode=0, bn=256, cta_group=2, max_ctas=mc, stats_out=st)
        st = st.sum(1)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t() + bias + res.float()
        all_ok &= report(f"PAIR bf16+bias+res in place", x, ref, 1e-2)
        all_ok &= report(f"   stats_out sum", st[:, 0], ref.sum(1), 2e-3)
        all_ok &= report(f"   stats_out sumsq", st[:, 1], (ref ** 2).sum(1), 2e-3)
    # raw stats consumed by an LN-folded GEMM (single CTA and pair)
    for cg in (1, 2):
        M, N, K = 2048, 1536, 512
        x = (torch.randn(M, K, device=dev) * 2 + 0.5).bfloat16()
        xf = x.float()
        raw = torch.stack([xf[:, :200].sum(1), (xf[:, :200] ** 2).sum(1), xf[:, 200:].sum(1), (xf[:, 200:] ** 2).sum(1)], 1).contiguous()
        gamma = torch.rand(K, device=dev) + 0.5; beta = torch.randn(K, device=dev) * 0.1
        W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev) * 0.1
        Wf = (W * gamma).bfloat16(); colsum = Wf.float().sum(1); biasf = b + W @ beta
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        run_gemm(x, Wf, out, bias=biasf, colsum=colsum, stats=raw, stats_raw=2, out_mode=0, bn=256, cta_group=cg)
        torch.cuda.synchronize()
        all_ok &= report(f"LN-fold from raw sums cta_group={cg}", out, torch.nn.functional.layer_norm(xf, (K,), gamma, beta, 1e-5) @ W.t() + b, 2e-2)
    # SwiGLU on pairs
    M, K, hid = 1024, 512, 1408
    a, _ = mk(M, 8, K)
    Wg = (torch.randn(hid, K, device=dev) / K ** 0.5); Wv = (torch.randn(hid, K, device=dev) / K ** 0.5)
    bg = torch.randn(hid, device=dev) * 0.1; bv = torch.randn(hid, device=dev) * 0.1
    Wp = torch.empty(2 * hid, K, device=dev); bp = torch.empty(2 * hid, device=dev)
    for t in range(hid // 128):
        Wp[t * 256:t * 256 + 128] = Wg[t * 128:(t + 1) * 128]; Wp[t * 256 + 128:(t + 1) * 256] = Wv[t * 128:(t + 1) * 128]
        bp[t * 256:t * 256 + 128] = bg[t * 128:(t + 1) * 128]; bp[t * 256 + 128:(t + 1) * 256] = bv[t * 128:(t + 1) * 128]
    Wp = Wp.bfloat16()
    out = torch.empty(M, hid, device=dev, dtype=torch.bfloat16)
    run_gemm(a, Wp, out, bias=bp, out_mode=0, swiglu=1, cta_group=2)
    torch.cuda.synchronize()
    g = a.float() @ Wg.bfloat16().float().t() + bg; v = a.float() @ Wv.bfloat16().float().t() + bv
    all_ok &= report("PAIR SwiGLU", out, torch.nn.functional.silu(g) * v, 2e-2)

    # 7. timing at the bench shapes (B=256 images -> M=262144)
    M = 262144
    for (name, N, K, bn, kw) in [("qkv", 1536, 512, 256, {}), ("out+res", 512, 512, 256, {"res": True}),
                                  ("w12 swiglu", 2816, 512, 256, {"swiglu": 1}), ("w3+res", 512, 1408, 256, {"res": True}),
                                  ("out+res bn128", 512, 512, 128, {"res": True})]:
        a, w = mk(M, N, K)
        nout = N // 2 if kw.get("swiglu") else N
        out = torch.empty(M, nout, device=dev, dtype=torch.bfloat16)
        res = torch.randn(M, nout, device=dev).bfloat16() if kw.get("res") else None
        bias = torch.randn(N, device=dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        iters = 10
        for cg in ((1, 2) if bn == 256 else (1,)):
            for _ in range(3):
                run_gemm(a, w, out, bias=bias, res=res, out_mode=0, swiglu=kw.get("swiglu", 0), bn=bn, cta_group=cg)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                run_gemm(a, w, out, bias=bias, res=res, out_mode=0, swiglu=kw.get("swiglu", 0), bn=bn, cta_group=cg)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            flops = 2.0 * M * N * K
            bytes_ = 2.0 * (M * K + N * K + M * nout * (2 if res is not None else 1))
            log(f"[perf] {name}: M{M} N{N} K{K} bn{bn} cta_group={cg}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {bytes_ / ms / 1e6:.0f} GB/s")
        # cuBLAS reference time for context (plain matmul, no epilogue)
        wt = w.t().contiguous()
        for _ in range(2):
            torch.matmul(a, wt)
        torch.cuda.synchronize(); e0.record()
        for _ in range(iters):
            torch.matmul(a, wt)
        e1.record(); torch.cuda.synchronize()
        log(f"        torch.matmul (cuBLAS) same shape: {e0.elapsed_time(e1) / iters:.3f} ms")
    log("ALL OK" if all_ok else "SOME FAILED")
    return 0 if all_ok else 1


if __name__ == "__main__":
    try:
        rc = main()
    except Exception as e:  # noqa: BLE001
        log("EXCEPTION:", repr(e))
        rc = 2
    sys.exit(rc)
