#!/bin/bash
# A/B of the mbarrier wait flavours (pm_common.cuh, PM_MBAR_FLAVOR): rebuilds the library on the GPU box per flavour
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02_mbar_flavors.txt
: > $OUT
for fl in "$@"; do
  export PM_NVCC_EXTRA="$fl"
  echo "=== PM_NVCC_EXTRA='$fl'" | tee -a $OUT
  python -m paintmind_b200.build --force > /dev/null 2>&1 || { echo "build failed" | tee -a $OUT; continue; }
  PM_AB_TIMEOUT=90 timeout 300 python scripts/attn3_ab.py old w16:1 2>&1 | tee -a $OUT
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-maskgit --no-train > gpurun_out/fl_bench.json 2> gpurun_out/fl_bench.err
  python - <<PY | tee -a $OUT
import json
d = json.loads(open("gpurun_out/fl_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "clk", d["clocks"]["sm_mhz"], "attn ms", round(d["roofline"]["avg_launch_ms"], 4),
      "vq", round(d["vq_lookups_per_s"] / 1e6))
print("   " + "  ".join(f'{k["kernel"].split("_", 1)[0]}{k["kernel"].split("_")[2] if k["kernel"].startswith("gemm") else ""}:{k["ms_total"] / k["launches"]:.4f}' for k in d["kernels"][:6]))
PY
done
unset PM_NVCC_EXTRA
