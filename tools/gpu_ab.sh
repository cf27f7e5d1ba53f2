#!/bin/bash
# A/B: kernel tests + short bench under two env settings given as args ("VAR=a" "VAR=b")
mkdir -p gpurun_out
i=0
for setting in "$@"; do
  i=$((i+1))
  env $setting timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 200 -x > gpurun_out/ab_tests_$i.log 2>&1
  echo "[$setting] tests: $(tail -1 gpurun_out/ab_tests_$i.log)"
  env $setting timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/ab_table_$i.json > gpurun_out/ab_bench_$i.log 2>&1
  echo "[$setting] bench rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ab_bench_$i.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_bench_$i.log) $(grep -o '"achieved": [0-9.]*' gpurun_out/ab_bench_$i.log)"
done
