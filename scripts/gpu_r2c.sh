#!/bin/bash
# round 2, third session: state of the tree after the pre-scaled-query attention kernels (pm_attn3 / pm_attn4)
set -u
mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python scripts/attn3_ab.py old 1,3,1,0 w16:1 w16:0 w16:2 > gpurun_out/${TAG}_attn3_ab.txt 2>&1
cat gpurun_out/${TAG}_attn3_ab.txt
timeout 300 python scripts/attn_clocks.py > gpurun_out/${TAG}_attn_clocks.txt 2>&1
cat gpurun_out/${TAG}_attn_clocks.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-maskgit --no-train > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn4_kernel -s 2 -c 1 -o gpurun_out/${TAG}_attn4 -f \
    python scripts/bench_e2e_quick.py 256 > gpurun_out/${TAG}_ncu_attn4.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_attn4.log
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
