# A/B of the fast-path forms of pm_attn4 (PM_ATTN4_SPLIT = 0: all 64 scores at once; 1: two halves of 32, no local-memory traffic;
# (2 = 1 + next step's first half loaded before the P store: measured, a loss, removed): isolated (burst + parity, attn3_ab.py) and inside the timed step of bench.py;
# alternating order, one box.  usage: bash scripts/attn4_split_ab.sh "1 2 1 2"
LIST=${1:-"0 1 0 1"}
for s in $LIST; do
  export PM_ATTN4_SPLIT=$s
  echo "split $s isolated: $(python scripts/attn3_ab.py w16:1 2>&1 | tail -1)"
done
for s in $LIST; do
  export PM_ATTN4_SPLIT=$s
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-maskgit --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split $s in-step:', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms  attn', round(d['roofline']['avg_launch_ms'],4), 'ms  clocks', d['clocks']['sm_mhz'], 'parity ok', d['check']['parity_vs_reference']['ok'])"
done
