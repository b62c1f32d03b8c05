#!/bin/bash
# round 2: full GPU test suite, headline bench, small-batch L2-residency experiment
set -u
mkdir -p gpurun_out
TAG=${1:-r02a}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 2500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
for b in 16 32 64; do
  timeout 300 python bench.py --steps 40 --warmup 5 --batch $b --no-cpu-baseline --no-maskgit --no-train > gpurun_out/${TAG}_bench_b$b.json 2> gpurun_out/${TAG}_bench_b$b.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_b$b.json").read().strip().splitlines()[-1])
print("batch $b:", round(d["value"]), "img/s", d["ms_per_step"], "ms/step", d["clocks"])
PY
done
