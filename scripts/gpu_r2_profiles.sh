#!/bin/bash
# round-2 evidence: final bench line, ncu launch list of the same command, ncu --set full of the dominant kernels, membound table
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
# launch list of the bench command (skip the 3 warm-up steps: 3 x 88 launches + set-up)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-maskgit --no-train > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 2 -c 1 -o gpurun_out/${TAG}_attn -f \
    python scripts/bench_e2e_quick.py 256 > gpurun_out/${TAG}_ncu_attn.log 2>&1
PM_ATTN_IMPL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_kernel -s 2 -c 1 -o gpurun_out/${TAG}_attn2 -f \
    python scripts/bench_e2e_quick.py 256 > gpurun_out/${TAG}_ncu_attn2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_exact4_kernel -s 1 -c 1 -o gpurun_out/${TAG}_vq4 -f \
    python scripts/vq_latency.py > gpurun_out/${TAG}_ncu_vq4.log 2>&1
PM_VQ_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vq_main_kernel -s 1 -c 1 -o gpurun_out/${TAG}_vq1 -f \
    python scripts/vq_latency.py > gpurun_out/${TAG}_ncu_vq1.log 2>&1
for kn in layernorm_kernel patchify8_u8_kernel; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -o gpurun_out/${TAG}_$kn -f \
    python scripts/membound_bench.py > gpurun_out/${TAG}_ncu_$kn.log 2>&1
done
timeout 300 python scripts/membound_bench.py > gpurun_out/${TAG}_membound.txt 2>&1
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
