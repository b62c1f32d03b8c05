"""Quick e2e timing of encode+decode at batch B with a per-kernel breakdown (CUDA events)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paintmind_b200 as pm  # noqa: E402
from paintmind_b200 import ops  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
cfg = ver2cfg["vit-s-vqgan"]
model = pm.create_model(arch="vqgan", version="vit-s-vqgan", pretrained=False)
model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=0))
model = model.to(dev).eval()
x = (torch.rand(B, 3, 256, 256, device=dev) * 2 - 1)


def step():
    z, loss, idx = model.encode(x)
    return model.decode(z)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
iters = 5
e0.record()
for _ in range(iters):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"B={B}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} img/s  ({B / ms * 1e3 * 138.58e9 / 1e12:.0f} TFLOP/s algorithmic)")

# per-op breakdown: wrap ops.* with event timing
acc = {}
orig = {}


def wrap(name):
    f = getattr(ops, name)
    orig[name] = f

    def g(*a, **k):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); r = f(*a, **k); e.record()
        key = name
        if name == "gemm":
            key = f"gemm N{a[1].shape[0]} K{a[0].shape[1]}" + (" swiglu" if k.get("swiglu") else "") + (" res" if k.get("res") is not None else "") + (" ln" if k.get("stats") is not None else "")
        acc.setdefault(key, []).append((s, e))
        return r
    setattr(ops, name, g)


for n in ["gemm", "attention", "layernorm", "patchify8", "vq_forward", "vq_codebook_prep", "split_rows32", "vq_gather"]:
    wrap(n)
import paintmind_b200.engine as eng  # noqa: E402
step(); torch.cuda.synchronize()
tot = 0.0
rows = []
for k, evs in acc.items():
    t = sum(s.elapsed_time(e) for s, e in evs)
    rows.append((t, k, len(evs))); tot += t
for t, k, n in sorted(rows, reverse=True):
    print(f"  {k:40s} x{n:3d}  {t:8.3f} ms  {100 * t / tot:5.1f}%")
print(f"  total (sum of kernels) {tot:.2f} ms")
