#!/bin/bash
# pix2pix layer shapes under the tap-GEMM planner's A/B switches; logs under gpurun_out/
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; env "$@" WAVE_CASES=p2p timeout 200 python tools/bench_wave.py 2>&1 | grep -E "^N" ; }
{
run default A=1
run nosplit B200GAN_TAPSPLIT=0
run nodual B200GAN_DUAL=1
run nodual_nosplit B200GAN_DUAL=1 B200GAN_TAPSPLIT=0
run nocta2 B200GAN_CTA2=0
run nocta2_nodual_nosplit B200GAN_CTA2=0 B200GAN_DUAL=1 B200GAN_TAPSPLIT=0
run nopersist B200GAN_PERSIST=0
} > gpurun_out/p2p_sweep.log 2>&1
cat gpurun_out/p2p_sweep.log
