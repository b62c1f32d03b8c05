"""Isolated timing of the GEMM shapes of one transformer block at B = 256 (M = 262144) with their real epilogues."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops  # noqa: E402
from paintmind_b200.engine import pack_swiglu_w12  # noqa: E402

dev = torch.device("cuda:0")
M, D = 262144, 512
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, D, device=dev, generator=g).bfloat16()
stats = torch.empty(M, ops.stats_parts(D), 2, device=dev)
stats2 = torch.empty_like(stats)


def timeit(fn, flops, name):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:40s} {ms:7.3f} ms  {flops / ms / 1e9:7.0f} TFLOP/s")


# produce valid statistics of x first (to_out-like GEMM writing x with stats_out)
wo = (torch.randn(D, D, device=dev, generator=g) * 0.04).bfloat16()
bo = torch.zeros(D, device=dev)
ao = torch.randn(M, D, device=dev, generator=g).bfloat16()
timeit(lambda: ops.gemm(ao, wo, x, bias=bo, res=x, stats_out=stats), 2 * M * D * D, "to_out  N=512  K=512  +res +stats_out")
wqkv = (torch.randn(3 * D, D, device=dev, generator=g) * 0.04).bfloat16()
qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
cs = wqkv.float().sum(1).contiguous()
bq = torch.zeros(3 * D, device=dev)
timeit(lambda: ops.gemm(x, wqkv, qkv, bias=bq, colsum=cs, stats=stats, stats_raw=ops.stats_parts(D)), 2 * M * 3 * D * D, "qkv     N=1536 K=512  LN-fold")
w12 = torch.randn(2736, D, device=dev, generator=g) * 0.04
b12 = torch.zeros(2736, device=dev)
w12p, cs12, b12p, hp = pack_swiglu_w12(w12, b12, torch.ones(D, device=dev), torch.zeros(D, device=dev))
h = torch.empty(M, hp, device=dev, dtype=torch.bfloat16)
timeit(lambda: ops.gemm(x, w12p, h, bias=b12p, colsum=cs12, swiglu=True, stats=stats, stats_raw=ops.stats_parts(D)), 2 * M * 2736 * D, "w12     N=2816 K=512  LN-fold SwiGLU")
w3 = (torch.randn(D, hp, device=dev, generator=g) * 0.03).bfloat16()
timeit(lambda: ops.gemm(h, w3, x, bias=bo, res=x, stats_out=stats2), 2 * M * D * 1368, "w3      N=512  K=1408 +res +stats_out")
