#!/bin/bash
# round-2 final evidence: GPU tests, bench (both arms), ncu launch list of the bench command, ncu --set full of the kernels touched last
set -u
mkdir -p gpurun_out
TAG=${1:-r02f}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 600 gpurun_out/${TAG}_bench_reference.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1200 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python scripts/membound_bench.py > gpurun_out/${TAG}_membound.txt 2>&1
cat gpurun_out/${TAG}_membound.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-maskgit --no-train > gpurun_out/${TAG}_ncu_bench.log 2>&1
for kn in layernorm_reg_kernel ln_bwd_kernel patchify8_u8_kernel maskgit_sample_block_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -o gpurun_out/${TAG}_$kn -f \
      python scripts/membound_bench.py > gpurun_out/${TAG}_ncu_$kn.log 2>&1
done
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
