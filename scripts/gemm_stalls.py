"""MMA-issuer stall accounting of pm_gemm_bf16 at the bench shapes (uses the `debug` counters)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = 262144
for (name, N, K, kw) in [("qkv", 1536, 512, {}), ("out+res", 512, 512, {"res": True}), ("w12 swiglu", 2816, 512, {"swiglu": True}),
                         ("w3+res", 512, 1408, {"res": True})]:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    nout = N // 2 if kw.get("swiglu") else N
    out = torch.empty(M, nout, device=dev, dtype=torch.bfloat16)
    res = torch.randn(M, nout, device=dev).bfloat16() if kw.get("res") else None
    bias = torch.randn(N, device=dev)
    for cg in (1, 2):
        dbg = torch.zeros(148, 4, device=dev, dtype=torch.int64)
        for _ in range(3):
            ops.gemm(a, w, out, bias=bias, res=res, swiglu=kw.get("swiglu", False), bn=256, cta_group=cg, debug=dbg)
        torch.cuda.synchronize()
        d = dbg[dbg[:, 2] > 0].double()
        tot = d[:, 2].mean().item()
        tiles = (M // (128 * cg)) * (N // 256) / (148 // cg)
        print(f"{name:12s} cta_group={cg}: total {tot:9.0f} cyc  acc-wait {100 * d[:, 0].mean().item() / tot:5.1f}%  operand-wait {100 * d[:, 1].mean().item() / tot:5.1f}%"
              f"  issue+other {100 * (1 - (d[:, 0].mean().item() + d[:, 1].mean().item()) / tot):5.1f}%  cycles/tile {tot / tiles:7.0f}  (mma-only {K // 64 * 4 * 128})")
