import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import paintmind_b200 as pm
from paintmind_b200.stage1.quantize import VectorQuantizer
from paintmind_b200 import ops
dev = torch.device("cuda:0")
vq = VectorQuantizer(8192, 32).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
zl = torch.nn.functional.normalize(torch.randn(65536, 32, device=dev, generator=g), dim=-1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5): vq.quantize_2d(zl)
torch.cuda.synchronize()
e0.record()
for _ in range(100): vq.quantize_2d(zl)
e1.record(); torch.cuda.synchronize()
print(f"eager quantize_2d: {e0.elapsed_time(e1) * 10:.1f} us per call")
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    vq.quantize_2d(zl)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    r = vq.quantize_2d(zl)
for _ in range(5): gr.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(100): gr.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay (prep + main + fills): {e0.elapsed_time(e1) * 10:.1f} us per call -> {65536 / (e0.elapsed_time(e1) * 1e-5) / 1e9:.3f} G lookups/s")
ops.PROFILE = {}
vq.quantize_2d(zl); torch.cuda.synchronize()
for k, evs in ops.PROFILE.items():
    print(k, [round(a.elapsed_time(b) * 1e3, 1) for a, b in evs], "us")
