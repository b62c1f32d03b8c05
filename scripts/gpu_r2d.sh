#!/bin/bash
# attention experiments: hinted waits / split P.V commit / per-scheduler ping-pong (PM_ATTN4_VARIANT="emu,opt")
set -u
mkdir -p gpurun_out
TAG=${1:-r02d}
PM_AB_TIMEOUT=90 timeout 900 python scripts/attn3_ab.py "$@" > gpurun_out/${TAG}_attn4_opt.txt 2>&1
cat gpurun_out/${TAG}_attn4_opt.txt
