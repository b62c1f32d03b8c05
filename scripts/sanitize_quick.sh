#!/bin/bash
# memcheck + synccheck over every kernel family, racecheck over the row kernels' callers (after kernel changes; the full sweep is sanitize.sh)
set -u
TAG=${1:-r02q}
mkdir -p gpurun_out
run() {
  local name=$1; shift
  local out=gpurun_out/${TAG}_sanitizer_${name}.txt
  local t0=$SECONDS
  timeout 900 compute-sanitizer "$@" > $out 2>&1
  local rc=$?
  echo "$name rc=$rc $((SECONDS - t0))s | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out | tail -1) | $(grep -c 'SANITIZE SWEEP DONE' $out) sweep(s) completed"
}
S="python scripts/sanitize_ops.py"
run memcheck  --tool memcheck  --error-exitcode 1 $S
run synccheck --tool synccheck --error-exitcode 1 $S
run racecheck --tool racecheck --racecheck-report analysis --error-exitcode 1 $S stage2 train
