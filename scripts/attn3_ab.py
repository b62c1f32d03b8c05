"""Time the pre-scaled-query attention kernel (pm_attn3.cu) and its variants against the round-1 kernel at the BASELINE shape
(B = 256, H = 8, N = 1024, d = 64) and check each against fp32 softmax attention.
usage: python scripts/attn3_ab.py [old | emu,defer,pref,pv2 ...]      e.g.  old 1,0,1 2,0,1"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

CHILD = r"""
import sys, math, torch
sys.path.insert(0, %r)
from paintmind_b200 import ops
pre = %r
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, H, N = 256, 8, 1024
qkv = torch.randn(B, N, 3 * 512, device=dev)
if pre:
    qkv[..., :512] *= 0.125 * 1.4426950408889634
qkv = qkv.bfloat16()
o = torch.empty(B, N, 512, device=dev, dtype=torch.bfloat16)
q, k, v = qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:]
run = lambda: ops.attention(q, k, v, o, H, 0.125, prescaled=pre)
for _ in range(4):
    run()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
sc = math.log(2.0) if pre else 0.125
def ref_attn(q, k, v):
    Bq = q.shape[0]
    qf = q.float().view(Bq, N, H, 64).transpose(1, 2); kf = k.float().view(Bq, N, H, 64).transpose(1, 2); vf = v.float().view(Bq, N, H, 64).transpose(1, 2)
    return (torch.softmax(qf @ kf.transpose(-1, -2) * sc, dim=-1) @ vf).transpose(1, 2).reshape(Bq, N, H * 64)
err = max((o[:2].float() - ref_attn(q[:2], k[:2], v[:2])).abs().max().item(), (o[-1:].float() - ref_attn(q[-1:], k[-1:], v[-1:])).abs().max().item())
print(f"RESULT ms={best:.4f} tflops={4*B*H*N*N*64/best/1e9:.1f} err={err:.4f}")
"""


def main():
    confs = sys.argv[1:] or ["old", "1,3,1,0", "w16:1", "w16:0", "w16:2"]
    for c in confs:
        env = dict(os.environ)
        pre = c != "old"
        if pre and c.startswith("w16:"):          # the 16-softmax-warp kernel, "w16:<emu>"
            env["PM_ATTN_PRE"] = "4"
            env["PM_ATTN4_VARIANT"] = c[4:]
        elif pre:
            env["PM_ATTN_PRE"] = "3"
            env["PM_ATTN3_VARIANT"] = c
        try:
            r = subprocess.run([sys.executable, "-c", CHILD % (str(ROOT), pre)], env=env, capture_output=True, text=True, timeout=int(os.environ.get("PM_AB_TIMEOUT", "240")))
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            print(f"variant {c:8s} {out[0] if out else 'FAILED: ' + r.stderr[-800:]}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"variant {c:8s} TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
