#!/bin/bash
# GPU test pass: each stage in its own process, bounded by timeouts; logs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 200 -s > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
if [ "$1" != "nobench" ]; then
timeout 900 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/gemm_table.json > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
fi
grep -E "passed|failed|rc=" gpurun_out/t_gpu.log | tail -3; tail -3 gpurun_out/smoke.log; tail -4 gpurun_out/bench.log | cut -c1-1500
