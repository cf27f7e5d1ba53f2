#!/bin/bash
# kernel tests, then short benches under the env settings given as args (same box)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 200 -x 2>&1 | tail -n 15
i=0
for setting in "$@"; do
  i=$((i+1))
  env $setting timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/ab_table_$i.json > gpurun_out/ab_bench_$i.log 2>&1
  echo "[$setting] bench rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ab_bench_$i.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_bench_$i.log) $(grep -o '"achieved": [0-9.]*' gpurun_out/ab_bench_$i.log)"
done
