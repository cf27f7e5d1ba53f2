#!/bin/bash
# GPU pass: each stage in its own process, bounded by timeouts; logs under gpurun_out/
# usage: tools/gpu_tests.sh [nobench|benchonly] [pytest -k expression]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
if [ "$1" != "benchonly" ]; then
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 -s ${2:+-k "$2"} > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
fi
if [ "$1" != "nobench" ]; then
timeout 600 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/gemm_table.json > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
for w in iwgan64 vae32 cnn28 pix2pix256; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --profile-out gpurun_out/gemm_table_$w.json > gpurun_out/bench_$w.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench_$w.log
done
timeout 300 python tools/bench_membound.py --out gpurun_out/membound.csv > gpurun_out/membound.log 2>&1
echo "membound rc=$?" >> gpurun_out/membound.log
fi
grep -E "passed|failed|rc=" gpurun_out/t_gpu.log | tail -3; tail -2 gpurun_out/smoke.log; for f in gpurun_out/bench.log gpurun_out/bench_*.log; do tail -2 $f | cut -c1-700; done; tail -3 gpurun_out/membound.log
