#!/bin/bash
# first GPU contact: each test file in its own process, bounded by timeouts
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x -k "conv_family and 4x16x16x64x64" > gpurun_out/t_first.log 2>&1
echo "first rc=$?" >> gpurun_out/t_first.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 > gpurun_out/t_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/t_kernels.log
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 200 -s > gpurun_out/t_model.log 2>&1
echo "model rc=$?" >> gpurun_out/t_model.log
tail -5 gpurun_out/t_first.log gpurun_out/t_kernels.log gpurun_out/t_model.log
