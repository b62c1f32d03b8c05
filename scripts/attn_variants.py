"""Time (and check) every attention kernel variant: one subprocess per PM_ATTN_VARIANT value.
usage: python scripts/attn_variants.py [variant ...]   (default: all variants compiled into the library)"""
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

CHILD = r"""
import sys, torch
sys.path.insert(0, %r)
from paintmind_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, H, N = 256, 8, 1024
qkv = torch.randn(B, N, 3 * 512, device=dev).bfloat16()
o = torch.empty(B, N, 512, device=dev, dtype=torch.bfloat16)
q, k, v = qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:]
for _ in range(4):
    ops.attention(q, k, v, o, H, 0.125)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    for _ in range(10):
        ops.attention(q, k, v, o, H, 0.125)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
# correctness on 2 batch items against fp32 softmax attention
qf = q[:2].float().view(2, N, H, 64).transpose(1, 2)
kf = k[:2].float().view(2, N, H, 64).transpose(1, 2)
vf = v[:2].float().view(2, N, H, 64).transpose(1, 2)
ref = torch.softmax(qf @ kf.transpose(-1, -2) * 0.125, dim=-1) @ vf
ref = ref.transpose(1, 2).reshape(2, N, 512)
err = (o[:2].float() - ref).abs().max().item()
# ragged cross-attention shape (77 keys)
kv = torch.randn(4, 77, 2 * 1024, device=dev).bfloat16()
q2 = torch.randn(4, N, 1024, device=dev).bfloat16()
o2 = torch.empty(4, N, 1024, device=dev, dtype=torch.bfloat16)
ops.attention(q2, kv[..., :1024], kv[..., 1024:], o2, 16, 0.125)
qf = q2.float().view(4, N, 16, 64).transpose(1, 2)
kf = kv[..., :1024].float().view(4, 77, 16, 64).transpose(1, 2)
vf = kv[..., 1024:].float().view(4, 77, 16, 64).transpose(1, 2)
ref2 = (torch.softmax(qf @ kf.transpose(-1, -2) * 0.125, dim=-1) @ vf).transpose(1, 2).reshape(4, N, 1024)
err2 = (o2.float() - ref2).abs().max().item()
print(f"RESULT ms={best:.4f} tflops={4*B*H*N*N*64/best/1e9:.1f} err_self={err:.4f} err_cross77={err2:.4f}")
""" % str(ROOT)


def variants():
    src = (ROOT / "paintmind_b200" / "csrc" / "pm_attn.cu").read_text()
    return [",".join(x.strip() for x in m) for m in re.findall(r"PM_ATTN_V\((-?\d+),\s*(-?\d+),\s*(-?\d+),\s*(-?\d+),\s*(-?\d+)\)", src)]


def main():
    vs = sys.argv[1:] or variants()
    for v in vs:
        env = dict(os.environ, PM_ATTN_VARIANT=v)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(f"variant emu,split,defer,chain,ping={v:12s} {out[0] if out else 'FAILED: ' + r.stderr[-400:]}", flush=True)


if __name__ == "__main__":
    main()
