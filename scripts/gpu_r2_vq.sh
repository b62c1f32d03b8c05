#!/bin/bash
# round 2, VQ kernel: parity tests, then A/B timing of the 1x-MMA kernel against the round-1 four-term kernel
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stage1.py -x -q -m gpu -k "vq or encode_decode" 2>&1 | tail -15 > gpurun_out/r02_vq_pytest.txt
cat gpurun_out/r02_vq_pytest.txt
timeout 300 python scripts/vq_latency.py > gpurun_out/r02_vq_latency.txt 2>&1
PM_VQ_MODE=4 timeout 300 python scripts/vq_latency.py > gpurun_out/r02_vq_latency_mode4.txt 2>&1
echo "--- new"; cat gpurun_out/r02_vq_latency.txt; echo "--- round-1 kernel"; cat gpurun_out/r02_vq_latency_mode4.txt
