"""VectorQuantizer with the reference surface (stage1/quantize.py:9-44): n_e, e_dim, beta,
``embedding`` (nn.Embedding, N(0,1) init), forward(z) -> (z_q, loss, indices),
decode_from_indice(indices).  The arithmetic is csrc/pm_vq.cu."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class VectorQuantizer(nn.Module):
    def __init__(self, n_e, e_dim, beta=0.25):
        super().__init__()
        self.n_e = n_e
        self.e_dim = e_dim
        self.beta = beta
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.normal_()
        self._last_hist = None      # codebook usage of the last forward (int64 [n_e]); new in this build
        self._last_sse = None       # sum (z_q - z)^2 of the last forward (float64 [1])
        self._prep = None           # (key, en, packed): normalised codebook operands, rebuilt when the weight changes
        self._cand = None           # (key, cand_val, cand_idx): scratch of the split-codebook path (small M only)

    def invalidate(self):
        """Drop the cached normalised codebook.  Needed only after an in-place write that bypasses autograd's version
        counter (``embedding.weight.data.copy_(...)``): ``data_ptr`` and ``_version`` are the cache key."""
        self._prep = None

    def _codebook_operands(self):
        """(en fp32 [n_e, 32], packed bf16 [n_e, 64]) of the current codebook.  The reference re-normalises the codebook on
        every forward (quantize.py:21); the result only depends on the weight, so it is cached on (data_ptr, _version,
        device) of the parameter and rebuilt by the prep kernel when an optimizer step / load_state_dict bumps the version."""
        w = self.embedding.weight
        key = (w.data_ptr(), w._version, w.device, w.dtype)
        if self._prep is None or self._prep[0] != key:
            E = w.detach()
            if E.dtype != torch.float32:
                E = E.float()
            en, packed = ops.vq_codebook_prep(E.contiguous())
            self._prep = (key, en, packed)
        return self._prep[1], self._prep[2]

    def _load_from_state_dict(self, *args, **kwargs):
        self._prep = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._prep = None
        self._cand = None
        return super()._apply(fn, *args, **kwargs)

    def _check(self, t):
        if not t.is_cuda or not self.embedding.weight.is_cuda:
            raise RuntimeError("paintmind_b200.VectorQuantizer runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.e_dim != 32:
            raise RuntimeError("paintmind_b200 VQ kernels are built for e_dim = 32")

    @torch.no_grad()
    @ops.on_device_of
    def quantize_2d(self, z2d, want_split=False):
        """z2d: fp32 [M, 32] (row stride multiple of 4).  Returns dict(idx, zq, zq_split, sse, hist)."""
        self._check(z2d)
        M = z2d.shape[0]
        dev = z2d.device
        en, packed = self._codebook_operands()
        idx = torch.empty(M, device=dev, dtype=torch.int64)
        zq = torch.empty(M, self.e_dim, device=dev, dtype=torch.float32)
        zs = torch.empty(M, 2 * self.e_dim, device=dev, dtype=torch.bfloat16) if want_split else None
        # usage histogram (int64 [n_e]) and the squared-error sum (float64 [1]) share one zero-filled allocation: one memset
        acc = torch.zeros(self.n_e + 1, device=dev, dtype=torch.int64)
        hist, sse = acc[:self.n_e], acc[self.n_e:].view(torch.float64)
        cv = ci = None
        if ops.vq_splits(M, self.n_e) > 1:       # the codebook is only split over several CTAs when M alone cannot fill the GPU
            if self._cand is None or self._cand[0] != (M, dev):
                self._cand = ((M, dev), torch.empty(8, M, device=dev, dtype=torch.float32),
                              torch.empty(8, M, device=dev, dtype=torch.int32))
            cv, ci = self._cand[1], self._cand[2]
        ops.vq_forward(z2d, en, packed, idx=idx, zq=zq, zq_split=zs, sse=sse, hist=hist, cand_val=cv, cand_idx=ci)
        self._last_hist, self._last_sse = hist, sse
        return dict(idx=idx, zq=zq, zq_split=zs, sse=sse, hist=hist)

    @torch.no_grad()
    @ops.on_device_of
    def forward(self, z):
        self._check(z)
        shape = z.shape
        if z.numel() == 0:       # empty input: empty outputs and the reference's nan (mean over zero elements)
            return (torch.empty(shape, device=z.device), torch.full((), float("nan"), device=z.device),
                    torch.empty(shape[:-1], dtype=torch.int64, device=z.device))
        z2d = z.detach().reshape(-1, self.e_dim)
        if z2d.dtype != torch.float32:
            z2d = z2d.float()
        if z2d.stride(1) != 1 or z2d.stride(0) % 4 != 0 or z2d.data_ptr() % 16 != 0:
            z2d = z2d.contiguous()
        r = self.quantize_2d(z2d)
        # loss = beta * mean((zq - z)^2) + mean((zq - z)^2)   (quantize.py:33), mean over all elements
        loss = (r["sse"] * ((1.0 + self.beta) / z2d.numel())).to(torch.float32).reshape(())
        return r["zq"].reshape(shape), loss, r["idx"].reshape(shape[:-1])

    @torch.no_grad()
    @ops.on_device_of
    def decode_from_indice(self, indices):
        self._check(indices)
        if indices.numel() == 0:
            return torch.empty(*indices.shape, self.e_dim, device=indices.device)
        en, _ = self._codebook_operands()          # rows of the cached l2norm(E): gathering them IS l2norm(E[idx])
        idx = indices.reshape(-1).to(torch.int64).contiguous()
        out = torch.empty(idx.numel(), self.e_dim, device=idx.device, dtype=torch.float32)
        ops.vq_gather(idx, en, False, out, None)
        return out.reshape(*indices.shape, self.e_dim)
