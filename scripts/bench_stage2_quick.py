"""Quick timing of one MaskGIT step / generate() with a per-kernel breakdown."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paintmind_b200 as pm  # noqa: E402
from paintmind_b200 import ops  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0")
cfg1, cfg2 = ver2cfg["vit-s-vqgan"], ver2cfg["paintmindv1"]
pipe = pm.create_model(arch="pipeline", version="paintmindv1", pretrained=False)
sd = {("vqgan." + k): v for k, v in synthetic.make_vqgan_state_dict(cfg1, seed=0).items()}
sd.update(synthetic.make_stage2_state_dict(cfg2, cfg1, seed=1, context_dim=1024))
pipe.load_state_dict(sd, strict=True)
pipe = pipe.to(dev).eval()
text = torch.randn(B, 77, 1024, device=dev)


def gen():
    return pipe.generate(text, timesteps=T, temperature=1.0, topk=5, save_interval=T)   # decode only the first step


gen(); gen()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
iters = 3
for _ in range(iters):
    gen()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = B * T * 437.72e9
print(f"generate B={B} T={T}: {ms:.1f} ms  {B / ms * 1e3:.1f} img/s  {fl / ms / 1e9:.0f} TFLOP/s (transformer only)")
ops.PROFILE = {}
gen(); torch.cuda.synchronize()
rows = []
for k, evs in ops.PROFILE.items():
    rows.append((sum(s.elapsed_time(e) for s, e in evs), k, len(evs)))
tot = sum(r[0] for r in rows)
for t, k, n in sorted(rows, reverse=True)[:14]:
    print(f"  {str(k):60s} x{n:4d} {t:9.3f} ms {100 * t / tot:5.1f}%")
print(f"  total kernels {tot:.1f} ms")
