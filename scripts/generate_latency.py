"""MaskGIT generate() at small and large batch, eager launches vs CUDA-graph replay of the transformer forward."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paintmind_b200 as pm  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda:0")
cfg, cfg2 = ver2cfg["vit-s-vqgan"], ver2cfg["paintmindv1"]
pipe = pm.create_model(arch="pipeline", version="paintmindv1", pretrained=False)
sd = {("vqgan." + k): v for k, v in synthetic.make_vqgan_state_dict(cfg, seed=0).items()}
sd.update(synthetic.make_stage2_state_dict(cfg2, cfg, seed=1, context_dim=1024))
pipe.load_state_dict(sd, strict=True)
pipe = pipe.to(dev).eval()
for B in (1, 8, 64):
    g = torch.Generator(device=dev).manual_seed(B)
    text = torch.randn(B, 77, 1024, device=dev, generator=g)
    res = {}
    for mode in (False, True):
        pipe.cuda_graph = mode
        outs = None
        for rep in range(3):
            torch.manual_seed(1234)        # the sampling noise is keyed on the torch generator state
            torch.cuda.synchronize(); t0 = time.perf_counter()
            outs = pipe.generate(text, timesteps=12, temperature=1.0, topk=5, save_interval=12)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        res[mode] = (dt, outs[0])
    same = torch.equal(res[False][1], res[True][1])
    print(f"B={B:3d}: generate() eager {res[False][0]:8.2f} ms   CUDA graph {res[True][0]:8.2f} ms   identical images: {same}")
