"""Bring-up counters of vq_main_kernel (needs a build with PM_NVCC_EXTRA=-DPM_VQ_DEBUG)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import _lib, ops  # noqa: E402
from paintmind_b200.stage1.quantize import VectorQuantizer  # noqa: E402

dev = torch.device("cuda:0")
vq = VectorQuantizer(8192, 32).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
zl = torch.nn.functional.normalize(torch.randn(65536, 32, device=dev, generator=g), dim=-1)
lib = _lib.load()
out = (C.c_ulonglong * 16)()
vq.quantize_2d(zl)
torch.cuda.synchronize()
lib.pm_debug_vq_counters(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
vq.quantize_2d(zl)
e1.record()
torch.cuda.synchronize()
lib.pm_debug_vq_counters(out, 1)
names = ["events (lane)", "pushes", "flagged rows", "-", "resolve rounds (lane-sum)", "scan cycles (lane-sum)",
         "resolve cycles (lane-sum)", "brute cycles (lane-sum)", "issuer barrier waits", "-", "drain t_full wait (lane-sum)",
         "issuer 2xMMA issue", "issuer 2xcommit", "reduce chunk0 (lane-sum)", "ld wait after reduce (lane-sum)"]
print(f"one call: {e0.elapsed_time(e1) * 1e3:.1f} us")
rows, lanes_items, issuer_items = 65536, 256 * 256, 256 * 2
for n, v in zip(names, out):
    if n == "-":
        continue
    print(f"{n:32s} {v:>14d}   per row {v / rows:10.3f}   per issuer-item {v / issuer_items:10.1f}")
