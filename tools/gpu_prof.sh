#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tapgemm|wgrad_kernel|im2col|col2im" -c 16 -o gpurun_out/prof_kernels python tools/prof_kernels.py > gpurun_out/prof_ncu.log 2>&1
echo "ncu kernels rc=$?" >> gpurun_out/prof_ncu.log
timeout 300 python tools/prof_step.py > gpurun_out/step_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py > gpurun_out/step_ncu.log 2>&1
echo "ncu step rc=$?" >> gpurun_out/step_ncu.log
tail -3 gpurun_out/prof_ncu.log gpurun_out/step_ncu.log gpurun_out/step_plain.log
