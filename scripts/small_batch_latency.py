"""Latency of encode + decode at small batch: wall clock per call (host-launch-bound?) against the device time of the same
work replayed from a CUDA graph.  python scripts/small_batch_latency.py"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paintmind_b200 as pm  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda:0")
cfg = ver2cfg["vit-s-vqgan"]
model = pm.create_model(arch="vqgan", version="vit-s-vqgan", pretrained=False)
model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=0), strict=True)
model = model.to(dev).eval()
for B in (1, 4, 16, 64):
    x = synthetic.make_images(B, 256, seed=3).to(dev)

    def step():
        z, loss, idx = model.encode(x)
        return model.decode(z), idx

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    eager = (time.perf_counter() - t0) / n * 1e3
    line = f"B={B:3d}: eager {eager:7.3f} ms/call ({B / eager * 1e3:7.0f} img/s)"
    if hasattr(model, "graphed"):
        run = model.graphed(x)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            run()
        torch.cuda.synchronize()
        g = (time.perf_counter() - t0) / n * 1e3
        rec, idx = step()
        rec_g, loss_g, idx_g, _ = run()
        line += f"   CUDA graph {g:7.3f} ms/call ({B / g * 1e3:7.0f} img/s)  same result: {bool(torch.equal(rec, rec_g) and torch.equal(idx, idx_g))}"
    print(line)
