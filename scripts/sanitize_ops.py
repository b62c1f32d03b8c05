"""Small-shape sweep over every kernel family of libpaintmind_b200 for compute-sanitizer (scripts/sanitize.sh).

Shapes are sized so that each tool (memcheck / racecheck / synccheck / initcheck; 10-1000x slowdown) finishes in minutes,
while still reaching every code path: 1-CTA and CTA-pair GEMMs with every epilogue, ragged and full attention tiles
(forward — both kernels — and both backward launches), weight-gradient split-K, LayerNorm / SwiGLU / VQ forward and
backward, codebook split + finalize, the MaskGIT sampling / re-mask / random-mask / cross-entropy kernels, uint8 ingest /
egress.  Results are also checked for finiteness so that a silent corruption shows up as a failure of this script.
usage: python scripts/sanitize_ops.py [stage1] [stage1_full] [stage2] [train]     (default: all)"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paintmind_b200 as pm  # noqa: E402
from paintmind_b200 import ops  # noqa: E402
from paintmind_b200.config import ver2cfg  # noqa: E402
from paintmind_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda:0")
what = set(sys.argv[1:]) or {"stage1", "stage1_full", "stage2", "train"}


def fin(*ts):
    for t in ts:
        assert torch.isfinite(t.float()).all(), "non-finite output"


def vqgan(name, seed):
    cfg = ver2cfg[name]
    m = pm.create_model(arch="vqgan", version=name, pretrained=False)
    m.load_state_dict(synthetic.make_vqgan_state_dict(cfg, seed=seed), strict=True)
    return cfg, m.to(dev).eval()


if "stage1" in what:
    cfg, m = vqgan("vit-tiny-test", 7)                       # 64-token sequences: ragged attention tiles, 1-CTA GEMMs, split VQ
    x = synthetic.make_images(3, cfg["enc"]["image_size"], seed=107).to(dev)
    z, loss, idx = m.encode(x)
    rec = m.decode(z)
    rec2 = m.decode_from_indice(idx)
    u8 = ((x.permute(0, 2, 3, 1) + 1) * 127.5).clamp(0, 255).to(torch.uint8).contiguous()
    z8, _, _ = m.encode_pixels(u8)
    px = m.decode_pixels(z8)
    fin(z, loss, rec, rec2, z8)
    print("stage1 tiny ok", float(loss))

if "stage1_full" in what:
    cfg, m = vqgan("vit-s-vqgan", 0)                         # 1024-token sequences: CTA-pair GEMMs, full attention tiles
    x = synthetic.make_images(2, 256, seed=100).to(dev)
    z, loss, idx = m.encode(x)
    rec = m.decode(z)
    fin(z, loss, rec)
    zl = F.normalize(torch.randn(600, 32, device=dev), dim=-1)   # VQ: one split, ragged row tile
    r = m.quantize.quantize_2d(zl)
    fin(r["zq"])
    print("stage1 vit-s ok", float(loss))

if "train" in what:
    cfg, m = vqgan("vit-tiny-test", 7)
    m.train()
    x = synthetic.make_images(2, cfg["enc"]["image_size"], seed=107).to(dev)
    rec, closs = m(x)
    (closs + F.l1_loss(rec, x) + F.mse_loss(rec, x)).backward()
    fin(*[p.grad for p in m.parameters()])
    cfg, m = vqgan("vit-s-vqgan", 0)                         # full-size tiles of the backward kernels, one image
    m.train()
    x = synthetic.make_images(1, 256, seed=100).to(dev)
    rec, closs = m(x)
    (closs + F.l1_loss(rec, x) + F.mse_loss(rec, x)).backward()
    fin(*[p.grad for p in m.parameters()])
    print("train ok")

if "stage2" in what:
    g = torch.Generator().manual_seed(3)
    V, Mrows = 8192, 96
    logits = torch.randn(Mrows, V, generator=g).to(dev)
    ids = torch.full((Mrows,), V, dtype=torch.int64, device=dev)
    ids[::3] = 5
    pred = torch.empty(Mrows, dtype=torch.int64, device=dev)
    sc = torch.empty(Mrows, device=dev)
    ops.maskgit_sample(logits, topk=5, temperature=0.7, ids=ids, pred_ids=pred, scores=sc, mask_id=V, seed=1, offset=2)
    ops.maskgit_sample(logits[:, :1000].contiguous(), topk=3, temperature=1.0, ids=None, pred_ids=pred, scores=sc, mask_id=V, seed=1, offset=3)
    ids2 = ids.view(3, 32).clone()
    ops.maskgit_remask(sc.view(3, 32), ids2, 7, V)
    z = torch.randn(3 * 32, 32, device=dev)
    mask = torch.empty(3, 32, device=dev)
    xo = torch.empty(3 * 32, 32, device=dev)
    ops.maskgit_random_mask(z, torch.zeros(32, device=dev), 3, 32, 8, mask=mask, x_out=xo, seed=4, offset=1)
    row_loss = torch.empty(Mrows, device=dev)
    out = torch.empty((), device=dev)
    sums = torch.empty(2, device=dev, dtype=torch.float64)
    ops.ce_label_smooth(logits, torch.randint(0, V, (Mrows,), generator=g).to(dev), mask.view(-1), 0.1, row_loss=row_loss, loss_out=out, sums_out=sums)
    fin(sc.clamp(min=-10), xo, out)
    # cross-attention over 77 keys (ragged key tile) and the 1024-d stage-2 GEMM shapes, through the public module
    from paintmind_b200.stage2 import CondTransformer
    tr = CondTransformer(32, 128, 64, 64, 256, 2, 1, 0.0, 96, 512).to(dev).eval()
    lg = tr(torch.randn(2, 64, 32, device=dev), torch.randn(2, 77, 96, device=dev))
    fin(lg)
    print("stage2 ok")
torch.cuda.synchronize()
print("SANITIZE SWEEP DONE")
