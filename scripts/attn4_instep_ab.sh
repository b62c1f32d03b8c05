# in-step A/B of the pm_attn4 exponential split (PM_ATTN4_VARIANT = quarter-pairs of every 4 on the FMA-pipe polynomial): alternating order, one box
for v in 1 0 2 1 0 2; do
  export PM_ATTN4_VARIANT=$v
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-maskgit --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('attn4 emu $v:', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms  attn in-step', round(d['roofline']['avg_launch_ms'],4), 'ms  clocks', d['clocks']['sm_mhz'])"
done
