#!/bin/bash
# A/B of the polling knobs of pm_attn4.cu (PM_A4_PROD_SLEEP / PM_A4_PV_SLEEP / PM_A4_QK_SLEEP): rebuilds on the box per flavour
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02_a4_sleep.txt
: > $OUT
for fl in "$@"; do
  export PM_NVCC_EXTRA="$fl"
  echo "=== PM_NVCC_EXTRA='$fl'" | tee -a $OUT
  python -m paintmind_b200.build --force > /dev/null 2>&1 || { echo "build failed" | tee -a $OUT; continue; }
  PM_AB_TIMEOUT=90 timeout 300 python scripts/attn3_ab.py w16:1 2>&1 | tee -a $OUT
  timeout 120 python scripts/attn_clocks.py 2>&1 | head -1 | tee -a $OUT
done
unset PM_NVCC_EXTRA
