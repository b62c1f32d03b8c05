"""GPU bring-up of the generator training step: VQModel.forward + backward against torch autograd of the fp32 torch
oracle on the same weights, images and (forced) code indices.  python scripts/bringup_train.py [version] [batch]"""
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paintmind_b200 as pm  # noqa: E402
from oracle import paintmind_oracle_torch as OT  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

version = sys.argv[1] if len(sys.argv) > 1 else "vit-s-vqgan"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
cfg = ver2cfg[version]
sd = synthetic.make_vqgan_state_dict(cfg, seed=0)
model = pm.create_model(arch="vqgan", version=version, pretrained=False)
model.load_state_dict(sd, strict=True)
model = model.to(dev).train()
img = synthetic.make_images(B, cfg["enc"]["image_size"], seed=11).to(dev)


def objective(rec, closs, img):
    return closs + F.l1_loss(rec.float(), img) + F.mse_loss(rec.float(), img)     # utils/trainer.py:207-215 minus LPIPS / GAN


rec, closs = model(img)
L = objective(rec, closs, img)
L.backward()
torch.cuda.synchronize()
ours = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
idx = model.train_engine().__dict__.get("_last_idx")
with torch.no_grad():
    _, _, idx = model.encode(img)

sdg = {k: v.to(dev).clone().requires_grad_(True) for k, v in sd.items()}
rec_o, closs_o, idx_o = OT.vqmodel_forward_train(img, sdg, cfg, idx=idx)
Lo = objective(rec_o, closs_o, img)
Lo.backward()
print(f"loss ours {L.item():.6f} oracle {Lo.item():.6f}; codebook loss {closs.item():.6f} / {closs_o.item():.6f}; "
      f"rec max err {(rec.float() - rec_o).abs().max().item():.4f}")
worst = []
for n, g in ours.items():
    go = sdg[n].grad
    if go is None:
        print("oracle has no grad for", n); continue
    rel = ((g - go).norm() / go.norm().clamp_min(1e-30)).item()
    cos = F.cosine_similarity(g.flatten(), go.flatten(), dim=0).item()
    worst.append((rel, cos, n, go.norm().item()))
worst.sort(reverse=True)
for rel, cos, n, nrm in worst[:25]:
    print(f"  rel_l2 {rel:.4f} cos {cos:.5f} |g| {nrm:.3e}  {n}")
print(f"median rel_l2 {sorted(w[0] for w in worst)[len(worst) // 2]:.4f}  max {worst[0][0]:.4f}  ({len(worst)} tensors)")
missing = [n for n, p in model.named_parameters() if p.grad is None]
print("params without grad:", missing)
if len(sys.argv) > 3:
    Bt = int(sys.argv[3])
    imgs = synthetic.make_images(8, cfg["enc"]["image_size"], seed=5).to(dev).repeat(Bt // 8, 1, 1, 1)
    for it in range(3):
        model.zero_grad(set_to_none=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rec, closs = model(imgs)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        objective(rec, closs, imgs).backward()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"B={Bt}: forward {1e3 * (t1 - t0):.1f} ms, backward {1e3 * (t2 - t1):.1f} ms -> {Bt / (t2 - t0):.0f} img/s, "
              f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
