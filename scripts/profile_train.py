"""Per-kernel time of one generator training step (forward + backward) at batch B: CUDA events around every launch
(ops.PROFILE).  python scripts/profile_train.py [B]"""
import sys
from collections import defaultdict
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paintmind_b200 as pm  # noqa: E402
from paintmind_b200 import ops  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
cfg = ver2cfg["vit-s-vqgan"]
model = pm.create_model(arch="vqgan", version="vit-s-vqgan", pretrained=False)
model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=0), strict=True)
model = model.to(dev).train()
img = synthetic.make_images(8, 256, seed=5).to(dev).repeat(B // 8, 1, 1, 1)


def step():
    model.zero_grad(set_to_none=True)
    rec, closs = model(img)
    (closs + F.l1_loss(rec, img) + F.mse_loss(rec, img)).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(3):
    step()
e.record(); torch.cuda.synchronize()
total = s.elapsed_time(e) / 3
print(f"B={B}: {total:.1f} ms/step = {B / total * 1e3:.0f} img/s (forward + backward, no profiling)")
ops.PROFILE = {}
step()
torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for key, evs in ops.PROFILE.items():
    name = key[0] if key[0] != "gemm" else f"gemm N={key[2]} K={key[3]}" + (" swiglu" if key[4] else "") + (" +res" if key[5] else "") + (" LN" if key[6] else "")
    if key[0] == "wgrad":
        name = f"wgrad N={key[2]} K={key[3]}"
    for a, b in evs:
        agg[name][0] += 1
        agg[name][1] += a.elapsed_time(b)
ops.PROFILE = None
tot = sum(v[1] for v in agg.values())
print(f"sum of profiled launches {tot:.1f} ms")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms:8.2f} ms  {100 * ms / tot:5.1f} %  x{n:<4d} {ms / n:7.3f} ms each  {name}")
