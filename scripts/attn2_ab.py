"""A/B of the two forward attention kernels (PM_ATTN_IMPL=1: pm_attn.cu, 2: pm_attn2.cu) and of the pm_attn2 variants:
one subprocess per configuration; prints time at the BASELINE shape (B = 256, H = 8, N = 1024) and errors against fp32
softmax attention (self-attention, ragged 77-key cross-attention, ragged query count, training outputs).
usage: python scripts/attn2_ab.py [impl[:variant] ...]      e.g.  1  2  2:1,0  2:2,8"""
import os
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

def ref_attn(q, k, v, heads, scale):
    Bq, Nq, _ = q.shape
    Nk = k.shape[1]
    qf = q.float().view(Bq, Nq, heads, 64).transpose(1, 2)
    kf = k.float().view(Bq, Nk, heads, 64).transpose(1, 2)
    vf = v.float().view(Bq, Nk, heads, 64).transpose(1, 2)
    s = qf @ kf.transpose(-1, -2) * scale
    return (torch.softmax(s, dim=-1) @ vf).transpose(1, 2).reshape(Bq, Nq, heads * 64), torch.logsumexp(s, dim=-1)

ref, _ = ref_attn(q[:2], k[:2], v[:2], H, 0.125)
err = (o[:2].float() - ref).abs().max().item()
ref, _ = ref_attn(q[-1:], k[-1:], v[-1:], H, 0.125)
err = max(err, (o[-1:].float() - ref).abs().max().item())
# ragged cross-attention shape (77 keys), 16 heads
kv = torch.randn(4, 77, 2 * 1024, device=dev).bfloat16()
q2 = torch.randn(4, N, 1024, device=dev).bfloat16()
o2 = torch.empty(4, N, 1024, device=dev, dtype=torch.bfloat16)
ops.attention(q2, kv[..., :1024], kv[..., 1024:], o2, 16, 0.125)
ref2, _ = ref_attn(q2, kv[..., :1024], kv[..., 1024:], 16, 0.125)
err2 = (o2.float() - ref2).abs().max().item()
# ragged query / key counts (200 tokens), large logits (scale 1.0: the lazy rescaling path fires), training outputs
q3 = (torch.randn(3, 200, 128, device=dev) * 2).bfloat16(); k3 = (torch.randn(3, 200, 128, device=dev) * 2).bfloat16(); v3 = torch.randn(3, 200, 128, device=dev).bfloat16()
o3 = torch.empty(3, 200, 128, device=dev, dtype=torch.bfloat16)
lse = ops.lse_buffer(3, 2, 200, dev); o32 = torch.empty(3, 200, 128, device=dev)
ops.attention_train(q3, k3, v3, o3, 2, 1.0, lse, o32)
ref3, lse3 = ref_attn(q3, k3, v3, 2, 1.0)
err3 = (o3.float() - ref3).abs().max().item()
err3f = (o32 - ref3).abs().max().item()
errl = (lse * 0.6931471805599453 - lse3).abs().max().item()
o3b = torch.empty_like(o3); ops.attention(q3, k3, v3, o3b, 2, 1.0)
err3 = max(err3, (o3b.float() - ref3).abs().max().item())
print(f"RESULT ms={best:.4f} tflops={4*B*H*N*N*64/best/1e9:.1f} err_self={err:.4f} err_cross77={err2:.4f} err_ragged200={err3:.4f} err_o32={err3f:.5f} err_lse={errl:.5f}")
""" % str(ROOT)


def main():
    confs = sys.argv[1:] or ["1", "2"]
    for c in confs:
        impl, _, var = c.partition(":")
        env = dict(os.environ, PM_ATTN_IMPL=impl)
        if var:
            env["PM_ATTN2_VARIANT" if impl == "2" else "PM_ATTN_VARIANT"] = var
        try:
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=240)
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            print(f"impl {c:10s} {out[0] if out else 'FAILED: ' + r.stderr[-600:]}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"impl {c:10s} TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
