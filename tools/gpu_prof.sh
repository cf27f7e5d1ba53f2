#!/bin/bash
# one ncu --set full capture of the heavy conv-family kernels (the launch list is tools/gpu_launchlist.sh:
# one ncu invocation per gpurun call)
mkdir -p gpurun_out
timeout 300 python tools/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tapgemm|wgrad|smallk|im2col|col2im|img_" -c 16 -o gpurun_out/prof_kernels python tools/prof_kernels.py > gpurun_out/prof_ncu.log 2>&1
echo "ncu kernels rc=$?" >> gpurun_out/prof_ncu.log
tail -n 3 gpurun_out/prof_ncu.log; tail -n 3 gpurun_out/prof_plain.log
