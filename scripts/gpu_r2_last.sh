#!/bin/bash
# last evidence run of round 2: GPU tests, both bench arms, ncu launch list of the bench command, ncu --set full of the dominant kernel,
# per-kernel profile of the training step
set -u
mkdir -p gpurun_out
TAG=${1:-r02k}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 400 gpurun_out/${TAG}_bench_reference.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-maskgit --no-train > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn4_kernel -s 2 -c 1 -o gpurun_out/${TAG}_attn4 -f \
    python scripts/bench_e2e_quick.py 256 > gpurun_out/${TAG}_ncu_attn4.log 2>&1
timeout 300 python scripts/profile_train.py > gpurun_out/${TAG}_train_profile.txt 2>&1
tail -45 gpurun_out/${TAG}_train_profile.txt
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
