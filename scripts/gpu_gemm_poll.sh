#!/bin/bash
# alternating A/B (thermal drift!) of the GEMM epilogue polling: all eight warps poll vs one; headline step only
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02_gemm_poll.txt
: > $OUT
for rep in 1 2 3; do
  for fl in "-DPM_GEMM_EPI_POLL_ALL=1" "-DPM_X=0"; do
    export PM_NVCC_EXTRA="$fl"
    python -m paintmind_b200.build --force > /dev/null 2>&1 || { echo "build failed" | tee -a $OUT; continue; }
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-maskgit --no-train > gpurun_out/fl_bench.json 2> gpurun_out/fl_bench.err
    python - <<PY | tee -a $OUT
import json
d = json.loads(open("gpurun_out/fl_bench.json").read().strip().splitlines()[-1])
print("$fl", "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "clk", d["clocks"]["sm_mhz"], "attn ms", round(d["roofline"]["avg_launch_ms"], 4),
      " ".join(f'{k["kernel"].split("_", 1)[0]}{k["kernel"].split("_")[2] if k["kernel"].startswith("gemm") else ""}:{k["ms_total"] / k["launches"]:.4f}' for k in d["kernels"][:5]))
PY
  done
done
unset PM_NVCC_EXTRA
