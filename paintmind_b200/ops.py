"""Tensor-level wrappers over the C-ABI (include/paintmind_b200.h).

Every function enqueues hand-written sm_100a kernels on torch's current CUDA stream and raises
RuntimeError on failure.  PyTorch only provides device memory and streams here.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import PM_OUT_BF16, PM_OUT_F32, PM_OUT_UNPATCH, PM_OUT_UNPATCH_U8  # noqa: F401


# Optional per-launch timing (bench.py): when PROFILE is a dict, every op records CUDA events on the
# launching stream around its kernel launch(es):  PROFILE[key] -> list of (start, end) events.
PROFILE = None
PROFILE_ONLY = None   # optional set of op names (key[0]): only these are bracketed by events (bench.py times the dominant kernel
                      # inside its timed region this way and the full per-kernel table in a separate, untimed pass)
LAUNCHES = 0          # kernels launched through this module (bench.py's gpu_launches)


# Optional NVTX ranges (PM_NVTX=1): phases of the engine (encode / quantize / decode / block i / MaskGIT step) and every op
# push a range on the calling thread, so that `ncu --nvtx --nvtx-include "pm.encode/"` or an nsys timeline can be cut by phase.
# Off by default: a range costs ~1 us of host time per op.
NVTX = os.environ.get("PM_NVTX", "0") not in ("", "0")


class nvtx_range:
    """with ops.nvtx_range("pm.encode"): ...   (no-op unless PM_NVTX=1)"""
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def nvtx_phase(name):
    """Decorator form of nvtx_range for engine entry points."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*a, **k):
            if not NVTX:
                return fn(*a, **k)
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapper
    return deco


def _prof_begin():
    if PROFILE is None or PROFILE_ONLY is not None:       # selective mode: only ops that call _prof_selected record events
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _prof_end(start, key, nkernels=1):
    global LAUNCHES
    LAUNCHES += nkernels
    if start is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        PROFILE.setdefault(key, []).append((start, ev))


def _prof_selected(name):
    """(start event or None) for an op that supports selective profiling: records only when `name` is selected."""
    if PROFILE is None or PROFILE_ONLY is None or name not in PROFILE_ONLY:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _prof_selected_end(start, key):
    if start is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        PROFILE.setdefault(key, []).append((start, ev))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("paintmind_b200 kernels need CUDA tensors (no CPU fallback)")
        dev = t.device.index if dev is None else dev
    if dev is not None and dev != torch.cuda.current_device():
        raise RuntimeError(f"paintmind_b200: tensors live on cuda:{dev} but the current device is cuda:{torch.cuda.current_device()}; "
                           "kernels launch on the current device's stream — wrap the call in `with torch.cuda.device(...)` "
                           "(the module-level entry points do this themselves)")


def on_device_of(fn):
    """Decorator for entry points: run with the first tensor argument's device current, so that kernels, tensor maps
    and per-device function attributes all refer to the device the data lives on (single-process multi-GPU use)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = next((a for a in args if torch.is_tensor(a)), None)
        if t is None or not t.is_cuda or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper


def gemm_tile_n(n, out_mode=PM_OUT_BF16, swiglu=False):
    """N-tile the library picks for bn = 0 (mirrors pm_gemm_bf16)."""
    if swiglu:
        return 256
    if out_mode in (PM_OUT_UNPATCH, PM_OUT_UNPATCH_U8):
        return 192 if n % 192 == 0 else 64
    if n % 256 == 0:
        return 256
    if n % 128 == 0:
        return 128
    if n % 64 == 0 or out_mode == PM_OUT_BF16:
        return 64
    return 32


def stats_parts(n, bn=None):
    """Partial (sum, sum-of-squares) pairs per row that a GEMM with N = n writes to `stats_out`."""
    bn = bn or gemm_tile_n(n)
    return 2 * ((n + bn - 1) // bn)


def gemm(a, w, out, *, bias=None, colsum=None, stats=None, pos=None, res=None, out_mode=PM_OUT_BF16,
         swiglu=False, bn=0, n=None, patch=0, channels=0, grid=0, max_ctas=0, stats_out=None, stats_raw=0,
         ln_eps=1e-5, cta_group=0, debug=None, res_mod=0):
    """out = epilogue(a[M,K] @ w[N,K]^T); see pm_gemm_bf16 in include/paintmind_b200.h."""
    _require_cuda(a, w, out)
    args = _lib.GemmArgs()
    args.a, args.w, args.out = _ptr(a), _ptr(w), _ptr(out)
    args.bias, args.colsum, args.stats = _ptr(bias), _ptr(colsum), _ptr(stats)
    args.pos, args.res = _ptr(pos), _ptr(res)
    args.lda, args.ldw = a.stride(0), w.stride(0)
    args.ld_out = out.stride(0) if out_mode not in (PM_OUT_UNPATCH, PM_OUT_UNPATCH_U8) else 0
    args.ld_pos = pos.stride(0) if pos is not None else 0
    args.ld_res = res.stride(0) if res is not None else 0
    args.M, args.N, args.K = a.shape[0], (n if n is not None else w.shape[0]), a.shape[1]
    args.pos_rows = pos.shape[0] if pos is not None else 0
    args.out_mode, args.swiglu, args.bn = out_mode, int(bool(swiglu)), bn
    args.patch, args.channels, args.grid, args.max_ctas = patch, channels, grid, max_ctas
    args.stats_raw, args.ln_eps, args.stats_out = int(stats_raw), float(ln_eps), _ptr(stats_out)
    args.cta_group = int(cta_group)
    args.debug = _ptr(debug)
    args.res_mod = int(res_mod)
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_gemm_bf16(C.byref(args), _stream()), "pm_gemm_bf16")
    _prof_end(t0, ("gemm", args.M, args.N, args.K, bool(swiglu), res is not None, stats is not None, out_mode))
    return out


def attention(q, k, v, o, heads, scale, prescaled=False):
    """q,o: [B, Nq, >=heads*64] views; k,v: [B, Nk, ...] views (bf16, last stride 1).  prescaled: q already carries
    scale * log2(e) (engine.py folds it into the to_q weight rows); `scale` is then ignored."""
    _require_cuda(q, k, v, o)
    args = _lib.AttnArgs()
    args.q, args.k, args.v, args.o = _ptr(q), _ptr(k), _ptr(v), _ptr(o)
    args.ldq, args.ldk, args.ldv, args.ldo = q.stride(1), k.stride(1), v.stride(1), o.stride(1)
    args.bsq, args.bsk, args.bsv, args.bso = q.stride(0), k.stride(0), v.stride(0), o.stride(0)
    args.B, args.H, args.Nq, args.Nk, args.head_dim = q.shape[0], heads, q.shape[1], k.shape[1], 64
    args.scale = float(scale)
    args.q_prescaled = int(bool(prescaled))
    t0 = _prof_begin()
    ts = _prof_selected("attention")
    _lib.check(_lib.load().pm_attn_fwd(C.byref(args), _stream()), "pm_attn_fwd")
    _prof_selected_end(ts, ("attention", args.B, args.H, args.Nq, args.Nk))
    _prof_end(t0, ("attention", args.B, args.H, args.Nq, args.Nk))
    return o


def vq_codebook_prep(E, en=None, packed=None):
    _require_cuda(E)
    n_e, e_dim = E.shape
    if en is None:
        en = torch.empty(n_e, e_dim, device=E.device, dtype=torch.float32)
    if packed is None:
        packed = torch.empty(n_e, 2 * e_dim, device=E.device, dtype=torch.bfloat16)
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_vq_codebook_prep(_ptr(E), n_e, e_dim, _ptr(en), _ptr(packed), _stream()),
               "pm_vq_codebook_prep")
    _prof_end(t0, ("vq_codebook_prep", n_e))
    return en, packed


def vq_splits(M, n_e):
    """Codebook splits pm_vq_fwd picks for splits = 0 (mirrors pm_vq_launch): > 1 only when the 256-row tiles alone cannot
    fill the SMs; the split path needs the cand_val / cand_idx scratch."""
    row_tiles = (M + 255) // 256
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    s = 1
    while row_tiles * s < sms and (n_e // (s * 2)) >= 8 * 128 and (n_e % (s * 2 * 128)) == 0 and s < 8:
        s *= 2
    return s


def vq_forward(z2d, en, packed, *, idx, zq=None, zq_split=None, sse=None, hist=None, cand_val=None, cand_idx=None,
               splits=0):
    _require_cuda(z2d, en, packed, idx)
    args = _lib.VqArgs()
    args.z, args.en, args.packed = _ptr(z2d), _ptr(en), _ptr(packed)
    args.cand_val, args.cand_idx = _ptr(cand_val), _ptr(cand_idx)
    args.idx, args.zq, args.zq_split = _ptr(idx), _ptr(zq), _ptr(zq_split)
    args.sse, args.hist = _ptr(sse), _ptr(hist)
    args.ldz = z2d.stride(0)
    args.M, args.n_e, args.e_dim, args.splits = z2d.shape[0], en.shape[0], en.shape[1], splits
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_vq_fwd(C.byref(args), _stream()), "pm_vq_fwd")
    _prof_end(t0, ("vq_forward", args.M, args.n_e), 2 if (splits != 1 and cand_val is not None and vq_splits(args.M, args.n_e) > 1) else 1)


def vq_gather(idx, table, normalize, out=None, out_split=None):
    _require_cuda(idx, table)
    M = idx.numel()
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_vq_gather(_ptr(idx), M, table.shape[0], table.shape[1], _ptr(table), int(normalize),
                                        _ptr(out), _ptr(out_split), _stream()), "pm_vq_gather")
    _prof_end(t0, ("vq_gather", M))


def split_rows32(src2d, out_split):
    _require_cuda(src2d, out_split)
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_split_rows32(_ptr(src2d), src2d.stride(0), src2d.shape[0], _ptr(out_split), _stream()),
               "pm_split_rows32")
    _prof_end(t0, ("split_rows32", src2d.shape[0]))
    return out_split


def patchify8(img, out):
    _require_cuda(img, out)
    B, Cc, H, W = img.shape
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_patchify8(_ptr(img), _ptr(out), B, Cc, H, W, _stream()), "pm_patchify8")
    _prof_end(t0, ("patchify8", B))
    return out


def patchify8_u8(img_u8, out):
    """uint8 NHWC [B, H, W, 3] pixels -> normalised bf16 patch rows (ToTensor + Normalize(0.5, 0.5) fused)."""
    _require_cuda(img_u8, out)
    B, H, W, Cc = img_u8.shape
    if Cc != 3 or img_u8.dtype != torch.uint8 or not img_u8.is_contiguous():
        raise RuntimeError("patchify8_u8 expects a contiguous uint8 [B, H, W, 3] tensor")
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_patchify8_u8(_ptr(img_u8), _ptr(out), B, H, W, _stream()), "pm_patchify8_u8")
    _prof_end(t0, ("patchify8_u8", B))
    return out


def layernorm(x, *, gamma=None, beta=None, y=None, stats=None, eps=1e-5):
    """y is None -> stats only; else y = LN(x) (and stats of y if given)."""
    _require_cuda(x)
    M, D = x.shape
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_layernorm(_ptr(x), x.stride(0), M, D, float(eps), _ptr(gamma), _ptr(beta), _ptr(y),
                                        y.stride(0) if y is not None else 0, _ptr(stats), _stream()), "pm_layernorm")
    _prof_end(t0, ("layernorm" if y is not None else "ln_stats", M, D))


def maskgit_sample(logits2d, *, topk, temperature, ids=None, pred_ids, scores, mask_id, noise=None, seed=0, offset=0,
                   step_tab=None, step_idx=None):
    """Fused top-k + gumbel arg-max + mask fill + confidence (generate.py:163-173) over fp32 logits [M, V].
    `step_tab` / `step_idx` (device tensors, see step_table): temperature / seed / offset are read on the device."""
    _require_cuda(logits2d, pred_ids, scores)
    a = _lib.MaskgitSampleArgs()
    a.step_tab, a.step_idx = _ptr(step_tab), _ptr(step_idx)
    a.logits, a.noise, a.ids, a.pred_ids, a.scores = _ptr(logits2d), _ptr(noise), _ptr(ids), _ptr(pred_ids), _ptr(scores)
    a.ld = logits2d.stride(0)
    a.ld_noise = noise.stride(0) if noise is not None else 0
    a.mask_id, a.seed, a.offset = int(mask_id), int(seed) & (2 ** 64 - 1), int(offset)
    a.M, a.V, a.topk, a.temperature = logits2d.shape[0], logits2d.shape[1], int(topk), float(temperature)
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_maskgit_sample(C.byref(a), _stream()), "pm_maskgit_sample")
    _prof_end(t0, ("maskgit_sample", a.M, a.V))


def maskgit_remask(scores2d, ids2d, k, mask_id):
    """ids.scatter(1, scores.topk(k).indices, mask_id) (generate.py:177-179), in place on ids."""
    _require_cuda(scores2d, ids2d)
    B, N = scores2d.shape
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_maskgit_remask(_ptr(scores2d), _ptr(ids2d), B, N, int(k), int(mask_id), _stream()),
               "pm_maskgit_remask")
    _prof_end(t0, ("maskgit_remask", B, N))


def step_table(temperatures, ks, seed, offsets, device):
    """Device table of pm_step_scalars (one 24-byte entry per MaskGIT step) as a uint8 tensor."""
    import numpy as np
    rec = np.zeros(len(ks), dtype=np.dtype([("temperature", "<f4"), ("k", "<i4"), ("seed", "<u8"), ("offset", "<u8")]))
    rec["temperature"], rec["k"], rec["seed"], rec["offset"] = temperatures, ks, np.uint64(int(seed) & (2 ** 64 - 1)), offsets
    return torch.from_numpy(rec.view(np.uint8).copy()).to(device)


def maskgit_remask_step(scores2d, ids2d, mask_id, step_tab, step_idx, ticket):
    """maskgit_remask with k read from step_tab[*step_idx]; advances *step_idx (last kernel of a graph-captured step)."""
    _require_cuda(scores2d, ids2d, step_tab, step_idx, ticket)
    B, N = scores2d.shape
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_maskgit_remask_step(_ptr(scores2d), _ptr(ids2d), B, N, int(mask_id), _ptr(step_tab), _ptr(step_idx),
                                                  _ptr(ticket), _stream()), "pm_maskgit_remask_step")
    _prof_end(t0, ("maskgit_remask", B, N))


def cast_bf16(src, dst):
    """fp32 -> bf16 copy (numel % 8 == 0)."""
    _require_cuda(src, dst)
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_cast_f32_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "pm_cast_f32_bf16")
    _prof_end(t0, ("cast_bf16", src.numel()))
    return dst


def maskgit_random_mask(z2d, mask_token, B, N, len_keep, *, mask, x_out=None, noise=None, seed=0, offset=0):
    """Pipeline.random_masking (generate.py:78-108): mask [B, N] fp32 (1 = replaced) and the masked token rows."""
    _require_cuda(mask)
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_maskgit_random_mask(_ptr(z2d), z2d.stride(0) if z2d is not None else 0, _ptr(noise),
                                                  int(seed) & (2 ** 64 - 1), int(offset), _ptr(mask_token), B, N,
                                                  int(len_keep), _ptr(mask), _ptr(x_out), _stream()),
               "pm_maskgit_random_mask")
    _prof_end(t0, ("maskgit_random_mask", B, N))


def ce_label_smooth(logits2d, label, mask, label_smoothing, *, row_loss, loss_out=None, sums_out=None):
    """Masked label-smoothed cross entropy (generate.py:110-123) over fp32 logits [M, V]."""
    _require_cuda(logits2d, label, row_loss)
    M, V = logits2d.shape
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_ce_label_smooth(_ptr(logits2d), logits2d.stride(0), M, V, _ptr(label), _ptr(mask),
                                              float(label_smoothing), _ptr(row_loss), _ptr(loss_out), _ptr(sums_out),
                                              _stream()), "pm_ce_label_smooth")
    _prof_end(t0, ("ce_label_smooth", M, V))


# ------------------------------------------------------------------------------------------------
# generator backward path
# ------------------------------------------------------------------------------------------------
_WORK = {}


def _workspace(name, nfloats, device):
    """fp32 scratch reused across calls (split-K partial tiles, column-sum partials).  One buffer per (purpose, device, STREAM):
    calls on one stream are ordered, and two streams never share a buffer, so concurrent use from several streams is safe
    (round 1 keyed on (purpose, device) only).  The C-ABI itself takes every workspace as an argument."""
    key = (name, str(device), torch.cuda.current_stream(device).cuda_stream)
    t = _WORK.get(key)
    if t is None or t.numel() < nfloats:
        t = torch.empty(max(int(nfloats), 1), device=device, dtype=torch.float32)
        _WORK[key] = t
    return t


def attention_train(q, k, v, o, heads, scale, lse, o32=None, prescaled=False):
    """attention() that also writes the base-2 log-sum-exp rows [B, heads, Nq] needed by attention_bwd and, optionally,
    an fp32 copy of the output ([B, Nq, heads*64] contiguous) for an accurate softmax-backward row term.  prescaled: q
    carries scale * log2(e) (see attention()); the matching backward call is attention_bwd(..., scale=ln 2) on the same q."""
    _require_cuda(q, k, v, o, lse, o32)
    args = _lib.AttnArgs()
    args.q, args.k, args.v, args.o = _ptr(q), _ptr(k), _ptr(v), _ptr(o)
    args.ldq, args.ldk, args.ldv, args.ldo = q.stride(1), k.stride(1), v.stride(1), o.stride(1)
    args.bsq, args.bsk, args.bsv, args.bso = q.stride(0), k.stride(0), v.stride(0), o.stride(0)
    args.B, args.H, args.Nq, args.Nk, args.head_dim = q.shape[0], heads, q.shape[1], k.shape[1], 64
    args.scale = float(scale)
    args.q_prescaled = int(bool(prescaled))
    args.lse, args.lse_ld = _ptr(lse), lse.stride(1)
    if o32 is not None:
        if o32.dtype != torch.float32 or not o32.is_contiguous():
            raise RuntimeError("attention_train: o32 must be a contiguous fp32 [B, Nq, heads*64] tensor")
        args.o32, args.ldo32 = _ptr(o32), o32.stride(1)
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_attn_fwd(C.byref(args), _stream()), "pm_attn_fwd")
    _prof_end(t0, ("attention", args.B, args.H, args.Nq, args.Nk))
    return o


def lse_buffer(B, heads, Nq, device):
    """[B, heads, Nq] fp32 view whose rows are padded to a multiple of 128 (attention_bwd copies whole 512-byte tiles)."""
    pad = (Nq + 127) // 128 * 128
    return torch.empty(B, heads, pad, device=device, dtype=torch.float32)[:, :, :Nq]


def attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, heads, scale, debug=None):
    """dq, dk, dv of softmax(scale q k^T) v; all [B, N, >= heads*64] bf16 views with last stride 1."""
    _require_cuda(q, k, v, o, d_o, lse, dq, dk, dv)
    B, Nq, Nk = q.shape[0], q.shape[1], k.shape[1]
    if lse.stride(1) % 128 != 0 or lse.stride(1) < Nq:
        raise RuntimeError("attention_bwd: lse rows must be padded to a multiple of 128 (see lse_buffer)")
    delta = _workspace("attn_delta", 2 * B * heads * lse.stride(1), q.device)
    a = _lib.AttnBwdArgs()
    a.q, a.k, a.v, a.o, a.d_o = _ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(d_o)
    a.lse, a.delta, a.dq, a.dk, a.dv = _ptr(lse), _ptr(delta), _ptr(dq), _ptr(dk), _ptr(dv)
    a.ldq, a.ldk, a.ldv, a.ldo, a.lddo = q.stride(1), k.stride(1), v.stride(1), o.stride(1), d_o.stride(1)
    a.lddq, a.lddk, a.lddv = dq.stride(1), dk.stride(1), dv.stride(1)
    a.bsq, a.bsk, a.bsv, a.bso, a.bsdo = q.stride(0), k.stride(0), v.stride(0), o.stride(0), d_o.stride(0)
    a.bsdq, a.bsdk, a.bsdv = dq.stride(0), dk.stride(0), dv.stride(0)
    a.B, a.H, a.Nq, a.Nk, a.head_dim = B, heads, Nq, Nk, 64
    a.scale = float(scale)
    a.o_is_f32 = int(o.dtype == torch.float32)
    a.lse_ld = lse.stride(1)
    a.debug = _ptr(debug)
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_attn_bwd(C.byref(a), _stream()), "pm_attn_bwd")
    _prof_end(t0, ("attention_bwd", B, heads, Nq, Nk), 3)


def wgrad(dy, x, out, n=None, k=None, accumulate=False):
    """out[N, K] (+)= dy[M, :N]^T @ x[M, :K]  (fp32 out; dy, x bf16 token-major)."""
    _require_cuda(dy, x, out)
    M = dy.shape[0]
    N = n if n is not None else dy.shape[1]
    K = k if k is not None else x.shape[1]
    lib = _lib.load()
    work = _workspace("wgrad", lib.pm_wgrad_workspace_floats(M, N, K), dy.device)
    t0 = _prof_begin()
    _lib.check(lib.pm_wgrad_bf16(_ptr(dy), dy.stride(0), _ptr(x), x.stride(0), M, N, K, _ptr(work), _ptr(out), out.stride(0),
                                 int(bool(accumulate)), _stream()), "pm_wgrad_bf16")
    _prof_end(t0, ("wgrad", M, N, K), 2)
    return out


def colsum(x, out, accumulate=False):
    """out[N] (+)= sum over rows of bf16 x[M, N]."""
    _require_cuda(x, out)
    M, N = x.shape
    lib = _lib.load()
    work = _workspace("colsum", lib.pm_colsum_workspace_floats(M, N), x.device)
    t0 = _prof_begin()
    _lib.check(lib.pm_colsum_bf16(_ptr(x), x.stride(0), M, N, _ptr(work), _ptr(out), int(bool(accumulate)), _stream()), "pm_colsum_bf16")
    _prof_end(t0, ("colsum", M, N), 2)
    return out


def layernorm_bwd(dn, x, gamma, dx, dgamma_dbeta, dres=None, eps=1e-5):
    """dx = LayerNorm'(x)^T dn (+ dres); dgamma_dbeta [2, D] fp32 = (sum dn * xhat, sum dn)."""
    _require_cuda(dn, x, gamma, dx, dgamma_dbeta)
    M, D = x.shape
    lib = _lib.load()
    work = _workspace("ln_bwd", lib.pm_layernorm_bwd_workspace_floats(M, D), x.device)
    t0 = _prof_begin()
    _lib.check(lib.pm_layernorm_bwd(_ptr(dn), dn.stride(0), _ptr(x), x.stride(0), _ptr(gamma), _ptr(dres),
                                    dres.stride(0) if dres is not None else 0, _ptr(dx), dx.stride(0), M, D, float(eps),
                                    _ptr(work), _ptr(dgamma_dbeta), _stream()), "pm_layernorm_bwd")
    _prof_end(t0, ("layernorm_bwd", M, D), 2)
    return dx


def swiglu_bwd(x12, dh, h, d12, b12=None):
    """d12 (and h) from x12, dh; b12 (fp32 [2 hp], optional) = column sums of d12 in the packed order, from the same pass."""
    _require_cuda(x12, dh, d12, b12)
    M, hp = dh.shape
    lib = _lib.load()
    work = _workspace("swiglu_bwd", lib.pm_swiglu_bwd_workspace_floats(M, hp), dh.device) if b12 is not None else None
    t0 = _prof_begin()
    _lib.check(lib.pm_swiglu_bwd(_ptr(x12), x12.stride(0), _ptr(dh), dh.stride(0), _ptr(h), h.stride(0) if h is not None else 0,
                                 _ptr(d12), d12.stride(0), M, hp, _ptr(work), _ptr(b12), _stream()), "pm_swiglu_bwd")
    _prof_end(t0, ("swiglu_bwd", M, hp), 2 if b12 is not None else 1)


def vq_bwd(z, idx, E, d_out, d_loss, beta, dz=None, dz_split=None, dE=None):
    _require_cuda(z, idx, E)
    M = z.shape[0]
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_vq_bwd(_ptr(z), z.stride(0), _ptr(idx), _ptr(E), E.shape[1], _ptr(d_out),
                                     d_out.stride(0) if d_out is not None else 0, _ptr(d_loss), float(beta), M,
                                     _ptr(dz), _ptr(dz_split), _ptr(dE), _stream()), "pm_vq_bwd")
    _prof_end(t0, ("vq_bwd", M))


def unpatchify8_bwd(d_img, rec, out):
    _require_cuda(d_img, rec, out)
    B, Cc, H, W = d_img.shape
    t0 = _prof_begin()
    _lib.check(_lib.load().pm_unpatchify8_bwd(_ptr(d_img), _ptr(rec), _ptr(out), B, Cc, H, W, _stream()), "pm_unpatchify8_bwd")
    _prof_end(t0, ("unpatchify8_bwd", B))
    return out
