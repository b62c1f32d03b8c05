import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 262144, 1536, 512
cg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); bias = torch.randn(N, device=dev)
cs = torch.randn(N, device=dev); st = torch.randn(M, 2, device=dev).abs() + 0.5
for _ in range(4):
    ops.gemm(a, w, out, bias=bias, colsum=cs, stats=st, bn=256, cta_group=cg)
torch.cuda.synchronize()
