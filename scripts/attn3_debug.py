"""Per-phase cycle accounting of attn3_kernel (needs a build with PM_NVCC_EXTRA=-DPM_A3_DEBUG): python scripts/attn3_debug.py"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, H, N = 256, 8, 1024
qkv = torch.randn(B, N, 3 * 512, device=dev)
qkv[..., :512] *= 0.125 * 1.4426950408889634
qkv = qkv.bfloat16()
o = torch.empty(B, N, 512, device=dev, dtype=torch.bfloat16)
q, k, v = qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:]
lib = _lib.load()
out = (C.c_ulonglong * 32)()
for _ in range(3):
    ops.attention(q, k, v, o, H, 0.125, prescaled=True)
torch.cuda.synchronize()
lib.pm_debug_a3_counters(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention(q, k, v, o, H, 0.125, prescaled=True)
e1.record()
torch.cuda.synchronize()
lib.pm_debug_a3_counters(out, 1)
steps = B * H * (N // 256) * (N // 128)          # two-tile steps over the whole grid
print(f"one call: {e0.elapsed_time(e1) * 1e3:.1f} us; {steps} steps, {steps / 148:.1f} per SM")
names = {0: "softmax t0: wait s_full", 1: "softmax t0: tcgen05.ld S", 2: "softmax t0: decide + s_free", 3: "softmax t0: exps + pv wait + st issue",
         4: "softmax t0: st wait + p_full", 5: "softmax t0: epilogue pv wait", 6: "softmax t0: epilogue", 7: "softmax t0: (pv wait inside 3)",
         8: "QK issuer: wait k_full (+q_full)", 9: "QK issuer: wait s_free (both tiles)", 10: "QK issuer: 10 MMAs", 11: "QK issuer: commits",
         12: "PV issuer(s): wait v_full", 13: "PV issuer(s): wait p_full", 14: "PV issuer(s): 16 MMAs", 15: "PV issuer(s): commits"}
names.update({24: "softmax (t0+t1): decide: max / sum check / rescale", 25: "softmax (t0+t1): decide: A write", 26: "softmax (t0+t1): tcgen05.fence before", 27: "softmax (t0+t1): exps chunk 0", 28: "softmax (t0+t1): exps chunk 1", 29: "softmax (t0+t1): exps chunks 2+3 (+ poll issue)", 30: "softmax (t0+t1): pv_done wait + fence", 31: "softmax (t0+t1): P stores"})
for i in range(8):
    names[16 + i] = names[i].replace("t0", "t1")
for i in sorted(names):
    print(f"{names[i]:44s} {out[i] / steps:9.1f} cycles per step")
