#!/bin/bash
# quick state check: GPU tests, headline bench, attention A/B, memory-bound table
set -u
mkdir -p gpurun_out
TAG=${1:-chk}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"],
      "attn ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3), "vq", round(d["vq_lookups_per_s"] / 1e6),
      "maskgit", round(d["maskgit"]["ms_per_generate"], 1), "train", round(d["train_step"]["ms_per_step"], 1))
for k in d["kernels"]: print("  ", k["kernel"], k["launches"], round(k["ms_total"] / k["launches"], 4), k["tflops"])
PY
PM_AB_TIMEOUT=90 timeout 400 python scripts/attn3_ab.py old w16:1 > gpurun_out/${TAG}_attn_ab.txt 2>&1
cat gpurun_out/${TAG}_attn_ab.txt
timeout 300 python scripts/membound_bench.py > gpurun_out/${TAG}_membound.txt 2>&1
cat gpurun_out/${TAG}_membound.txt
