import sys, torch, torch.nn.functional as F
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paintmind_b200 as pm
from paintmind_b200.config import ver2cfg
from paintmind_b200.utils import synthetic
dev = torch.device("cuda:0")
cfg = ver2cfg["vit-s-vqgan"]
model = pm.create_model(arch="vqgan", version="vit-s-vqgan", pretrained=False)
model.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=0), strict=True)
model = model.to(dev).train()
img = synthetic.make_images(8, 256, seed=5).to(dev).repeat(32, 1, 1, 1)
res = {}
for keep in (True, False):
    model.train_engine().keep_attention = keep
    torch.cuda.reset_peak_memory_stats()
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        rec, closs = model(img)
        (closs + F.l1_loss(rec, img) + F.mse_loss(rec, img)).backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        model.zero_grad(set_to_none=True)
        rec, closs = model(img)
        (closs + F.l1_loss(rec, img) + F.mse_loss(rec, img)).backward()
    e1.record(); torch.cuda.synchronize()
    res[keep] = {n: p.grad.clone() for n, p in model.named_parameters()}
    print(f"keep={keep}: {e0.elapsed_time(e1) / 3:.1f} ms/step, peak {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
worst = max(float((res[True][n] - res[False][n]).norm() / res[False][n].norm().clamp_min(1e-30)) for n in res[True])
print(f"kept vs recomputed gradients: worst relative L2 difference {worst:.2e}")
