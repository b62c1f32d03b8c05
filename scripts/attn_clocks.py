"""SM clock / power while a forward-attention kernel runs back to back (is the isolated kernel power-capped?).
usage: python scripts/attn_clocks.py [old|pre]   (PM_ATTN_PRE / PM_ATTN3_VARIANT / PM_ATTN4_VARIANT pick the pre-scaled kernel)"""
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from paintmind_b200 import ops  # noqa: E402

pre = (sys.argv[1] if len(sys.argv) > 1 else "pre") != "old"
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, H, N = 256, 8, 1024
qkv = torch.randn(B, N, 3 * 512, device=dev)
if pre:
    qkv[..., :512] *= 0.125 * 1.4426950408889634
qkv = qkv.bfloat16()
o = torch.empty(B, N, 512, device=dev, dtype=torch.bfloat16)
q, k, v = qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:]
samples = []
stop = False


def sampler():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True)
        samples.append(r.stdout.strip())
        time.sleep(0.05)


for _ in range(10):
    ops.attention(q, k, v, o, H, 0.125, prescaled=pre)
torch.cuda.synchronize()
th = threading.Thread(target=sampler)
th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3000
for _ in range(n):
    ops.attention(q, k, v, o, H, 0.125, prescaled=pre)
e1.record()
torch.cuda.synchronize()
stop = True
th.join()
ms = e0.elapsed_time(e1) / n
print(f"{'pre-scaled' if pre else 'old'}: {ms:.4f} ms per call over {n} calls = {4 * B * H * N * N * 64 / ms / 1e9:.0f} TFLOP/s")
mid = samples[len(samples) // 4: -1]
print("samples (sm MHz, W, sw_power_cap):", mid[:: max(1, len(mid) // 8)])
