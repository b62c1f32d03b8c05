#!/bin/bash
# One GPU session: tests, bench, launch list, ncu captures of the top kernels.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
if [ "${2:-full}" = "full" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-maskgit --no-train > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log
# 4th gemm launch of a step = layer-0 w12 (SwiGLU epilogue); 2nd = qkv (LN fold)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/${TAG}_swiglu -f \
    python scripts/bench_e2e_quick.py 256 > gpurun_out/${TAG}_ncu_swiglu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -o gpurun_out/${TAG}_qkv -f \
    python scripts/bench_e2e_quick.py 256 > gpurun_out/${TAG}_ncu_qkv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 2 -c 1 -o gpurun_out/${TAG}_attn -f \
    python scripts/bench_e2e_quick.py 256 > gpurun_out/${TAG}_ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vq_main_kernel -s 1 -c 1 -o gpurun_out/${TAG}_vq -f \
    python scripts/bench_e2e_quick.py 256 > gpurun_out/${TAG}_ncu_vq.log 2>&1
# memory-bound kernels: one capture each (DRAM bytes and throughput) out of the bandwidth micro-benchmark
for kn in ce_rows_kernel maskgit_sample_block_kernel layernorm_kernel patchify8_u8_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -o gpurun_out/${TAG}_$kn -f \
    python scripts/membound_bench.py > gpurun_out/${TAG}_ncu_$kn.log 2>&1
done
# generator backward path: per-kernel profile of a training step, kernel parity + timings, ncu of the two new tensor kernels
timeout 600 python scripts/profile_train.py 256 > gpurun_out/${TAG}_train_profile.txt 2>&1
timeout 600 python scripts/bringup_bwd.py > gpurun_out/${TAG}_bwd_kernels.txt 2>&1
timeout 300 python scripts/attn_bwd_stalls.py 256 > gpurun_out/${TAG}_attn_bwd_stalls.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 8 -c 2 -o gpurun_out/${TAG}_attn_bwd -f \
    python scripts/attn_bwd_stalls.py 256 > gpurun_out/${TAG}_ncu_attn_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 40 -c 1 -o gpurun_out/${TAG}_wgrad -f \
    python scripts/bringup_bwd.py full wgrad > gpurun_out/${TAG}_ncu_wgrad.log 2>&1
for kn in swiglu_bwd_kernel ln_bwd_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 6 -c 1 -o gpurun_out/${TAG}_$kn -f \
    python scripts/bringup_bwd.py full swiglu,ln > gpurun_out/${TAG}_ncu_$kn.log 2>&1
done
fi
ls -la gpurun_out/
