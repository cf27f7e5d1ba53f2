#!/bin/bash
# A/B over epilogue warp counts: epilogue micro-bench + kernel tests + short bench, all on one box
mkdir -p gpurun_out
for w in "$@"; do
  echo "=== B200GAN_EPI_WARPS=$w"
  B200GAN_EPI_WARPS=$w timeout 300 python tools/bench_epi.py 2>&1 | tee gpurun_out/epi_$w.log
  B200GAN_EPI_WARPS=$w timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 200 -x 2>&1 | tail -1
  B200GAN_EPI_WARPS=$w timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/epi_table_$w.json > gpurun_out/epi_bench_$w.log 2>&1
  echo "bench rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/epi_bench_$w.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/epi_bench_$w.log) $(grep -o '"achieved": [0-9.]*' gpurun_out/epi_bench_$w.log)"
done
