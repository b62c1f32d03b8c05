import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops
dev = torch.device("cuda:0")
B, H, N = 256, 8, 1024
qkv = torch.randn(B, N, 1536, device=dev).bfloat16()
o = torch.empty(B, N, 512, device=dev, dtype=torch.bfloat16)
o32 = torch.empty(B, N, 512, device=dev)
lse = ops.lse_buffer(B, H, N, dev)
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
q, k, v = qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:]
print(f"inference {t(lambda: ops.attention(q, k, v, o, H, 0.125)):.3f} ms;  +lse {t(lambda: ops.attention_train(q, k, v, o, H, 0.125, lse)):.3f} ms;  +lse +o32 {t(lambda: ops.attention_train(q, k, v, o, H, 0.125, lse, o32)):.3f} ms")
