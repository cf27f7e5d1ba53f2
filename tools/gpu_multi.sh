#!/bin/bash
# weak-scaling bench at N GPUs of one node (torchrun, NCCL)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1
echo "rc=$?" >> gpurun_out/bench_n$N.log
grep -E "metric|rc=" gpurun_out/bench_n$N.log | cut -c1-400
