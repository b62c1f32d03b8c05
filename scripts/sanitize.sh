#!/bin/bash
# compute-sanitizer pass over every kernel family.  Run on a B200 box:  bash scripts/sanitize.sh [tag]
# Output: gpurun_out/<tag>_sanitizer_<tool>[_variant].txt (full logs) and one summary line per run on stdout.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
run() {  # name, extra env, tool args..., then the sweep arguments after "--"
  local name=$1; shift
  local envs=$1; shift
  local out=gpurun_out/${TAG}_sanitizer_${name}.txt
  local t0=$SECONDS
  env $envs timeout 1500 compute-sanitizer "$@" > $out 2>&1
  local rc=$?
  echo "$name rc=$rc $((SECONDS - t0))s | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out | tail -1) | $(grep -c 'SANITIZE SWEEP DONE' $out) sweep(s) completed"
}
S="python scripts/sanitize_ops.py"
run memcheck   "PM_X=0" --tool memcheck  --error-exitcode 1 $S
run memcheck_attn2_vq1 "PM_ATTN_IMPL=2 PM_VQ_MODE=1" --tool memcheck --error-exitcode 1 $S stage1 stage1_full
run synccheck  "PM_X=0" --tool synccheck --error-exitcode 1 $S
run initcheck  "PM_X=0" --tool initcheck --error-exitcode 1 $S stage1 stage2 train
run racecheck  "PM_X=0" --tool racecheck --racecheck-report analysis --error-exitcode 1 $S stage1 stage2 train
run racecheck_full "PM_X=0" --tool racecheck --racecheck-report analysis --error-exitcode 1 $S stage1_full
