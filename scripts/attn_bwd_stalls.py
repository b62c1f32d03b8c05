"""Cycle accounting of the attention backward kernels (debug counters of pm_attn_bwd): python scripts/attn_bwd_stalls.py [B]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H, N, inner = 8, 1024, 512
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B, N, 3 * inner, device=dev, generator=g).bfloat16()
q, k, v = qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
o = torch.empty(B, N, inner, device=dev, dtype=torch.bfloat16)
lse = ops.lse_buffer(B, H, N, dev)
do = torch.randn(B, N, inner, device=dev, generator=g).bfloat16()
dqkv = torch.empty_like(qkv)
ops.attention_train(q, k, v, o, H, 0.125, lse)
dbg = torch.zeros(256, 8, device=dev, dtype=torch.int64)
args = (q, k, v, o, do, lse, dqkv[..., :inner], dqkv[..., inner:2 * inner], dqkv[..., 2 * inner:], H, 0.125)
for _ in range(3):
    ops.attention_bwd(*args)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    ops.attention_bwd(*args)
e.record(); torch.cuda.synchronize()
print(f"attn bwd B={B}: {s.elapsed_time(e) / 10:.3f} ms")
ops.attention_bwd(*args, debug=dbg)
torch.cuda.synchronize()
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
d = dbg[:n_sm].double().mean(0).tolist()
steps = B * H * (N // 128) * (N // 128) / n_sm
print(f"last kernel (dQ), mean cycles per CTA ({steps:.0f} steps each):")
print(f"  MMA thread : wait s_free {d[0] / steps:.0f}  wait C {d[1] / steps:.0f}  wait P {d[2] / steps:.0f}  total {d[3] / steps:.0f} per step")
print(f"  compute w0 : wait S {d[4] / steps:.0f}  wait acc {d[5] / steps:.0f}  math {d[6] / steps:.0f} per step;  epilogue {d[7] / (steps / (N // 128)):.0f} per item")
