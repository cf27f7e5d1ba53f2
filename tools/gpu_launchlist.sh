#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_step.py > gpurun_out/step_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py > gpurun_out/step_ncu.log 2>&1
echo "ncu step rc=$?" >> gpurun_out/step_ncu.log
tail -2 gpurun_out/step_ncu.log
