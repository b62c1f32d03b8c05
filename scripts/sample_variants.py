"""Time the MaskGIT sampling kernel variants (PM_MASKGIT_VARIANT) on [65536, 8192] fp32 logits, L2 flushed."""
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
M, V = 65536, 8192
g = torch.Generator(device=dev).manual_seed(0)
logits = torch.randn(M, V, device=dev, generator=g)
ids = torch.full((M,), V, device=dev, dtype=torch.int64)
pred = torch.empty(M, device=dev, dtype=torch.int64); sc = torch.empty(M, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def run():
    ops.maskgit_sample(logits, topk=5, temperature=1.0, ids=ids, pred_ids=pred, scores=sc, mask_id=V, seed=1, offset=1)
run(); run(); torch.cuda.synchronize()
ts = []
for _ in range(7):
    flush.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); ms = ts[len(ts) // 2]
print(f"RESULT ms={ms:.4f} GB/s={M * V * 4 / ms / 1e6:.0f} checksum={int(pred.sum())}")
""" % str(ROOT)
names = {"0": "auto (block-per-row staged)", "1": "warp-per-row staged", "2": "streaming insertion"}
for v in (sys.argv[1:] or ["0", "1", "2"]):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, PM_MASKGIT_VARIANT=v), capture_output=True, text=True, timeout=300)
    out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print(f"variant {v} {names.get(v, ''):28s} {out[0] if out else 'FAILED ' + r.stderr[-300:]}", flush=True)
