import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops
dev = torch.device("cuda:0")
M, V = 65536, 8192
logits = torch.randn(M, V, device=dev)
pred = torch.empty(M, device=dev, dtype=torch.int64); sc = torch.empty(M, device=dev)
ids = torch.full((M,), V, device=dev, dtype=torch.int64)
for _ in range(3):
    ops.maskgit_sample(logits, topk=5, temperature=1.0, ids=ids, pred_ids=pred, scores=sc, mask_id=V, seed=1, offset=1)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.maskgit_sample(logits, topk=5, temperature=1.0, ids=ids, pred_ids=pred, scores=sc, mask_id=V, seed=1, offset=1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"maskgit_sample {M}x{V}: {ms:.3f} ms  {M * V * 4 / ms / 1e6:.0f} GB/s")
