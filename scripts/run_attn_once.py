import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops
dev = torch.device("cuda:0")
B, H, N = 256, 8, 1024
qkv = torch.randn(B, N, 3 * 512, device=dev).bfloat16()
o = torch.empty(B, N, 512, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    ops.attention(qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:], o, H, 0.125)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention(qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:], o, H, 0.125)
e1.record(); torch.cuda.synchronize()
print(f"attention B256 H8 N1024: {e0.elapsed_time(e1) / 10:.3f} ms")
