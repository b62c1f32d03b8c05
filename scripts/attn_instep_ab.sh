for cfg in "1 1,1,3,8,0" "1 0,1,3,0,0" "1 2,1,3,8,0" "1 1,1,3,0,0" "2 1,8" "2 0,0"; do
  set -- $cfg
  if [ "$1" = "1" ]; then export PM_ATTN_IMPL=1 PM_ATTN_VARIANT=$2; unset PM_ATTN2_VARIANT; else export PM_ATTN_IMPL=2 PM_ATTN2_VARIANT=$2; unset PM_ATTN_VARIANT; fi
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-maskgit --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('impl $1 variant $2:', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms  attn', round(d['roofline']['avg_launch_ms'],4), 'ms  clocks', d['clocks']['sm_mhz'])"
done
