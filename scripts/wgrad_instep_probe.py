"""Why does the qkv weight gradient (wgrad N=1536 K=512) take 0.64 ms inside the training step and 0.39 ms alone?  Times the same
call (a) back to back, (b) after an attention backward, (c) after attention backward + LayerNorm recompute, (d) with dy = the
buffer the attention backward wrote (real gradients) instead of random data."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, H, N, inner, D = 256, 8, 1024, 512, 512
M = B * N
qkv = (torch.randn(B, N, 3 * inner, device=dev) * 0.5).to(torch.bfloat16)
q, k, v = qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
o = torch.empty(B, N, inner, device=dev, dtype=torch.bfloat16)
lse = ops.lse_buffer(B, H, N, dev)
ops.attention_train(q, k, v, o, H, 0.125, lse)
do = (torch.randn(B, N, inner, device=dev) * 0.01).to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
x = torch.randn(M, D, device=dev).to(torch.bfloat16)
nbuf = torch.empty_like(x)
g1 = torch.ones(D, device=dev); b1 = torch.zeros(D, device=dev)
dy_rand = (torch.randn(M, 3 * inner, device=dev) * 0.01).to(torch.bfloat16)
w = torch.empty(3 * inner, D, device=dev)


def bwd():
    ops.attention_bwd(q, k, v, o, do, lse, dqkv[..., :inner], dqkv[..., inner:2 * inner], dqkv[..., 2 * inner:], H, 0.125)


def timed(pre, dy, reps=10):
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for i in range(reps):
        pre()
        e0[i].record()
        ops.wgrad(dy, nbuf, w)
        e1[i].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
    return ts[len(ts) // 2]


bwd(); ops.layernorm(x, gamma=g1, beta=b1, y=nbuf); torch.cuda.synchronize()
d2 = dqkv.view(M, 3 * inner)
for name, pre, dy in (("alone, random dy", lambda: None, dy_rand),
                      ("alone, dy = attention-backward output", lambda: None, d2),
                      ("after attention backward, random dy", bwd, dy_rand),
                      ("after attention backward, its output", bwd, d2),
                      ("after attention backward + LayerNorm, its output", lambda: (bwd(), ops.layernorm(x, gamma=g1, beta=b1, y=nbuf)), d2)):
    ms = timed(pre, dy)
    print(f"wgrad N=1536 K=512 {name:55s}: {ms:.3f} ms  {2.0 * M * 1536 * 512 / ms / 1e9:.0f} TFLOP/s")
