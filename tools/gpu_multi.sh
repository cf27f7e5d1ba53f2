#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
export B200GAN_BENCH_VERBOSE=1 NCCL_DEBUG=WARN
B200GAN_CUDA_GRAPHS=0 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 --no-roofline > gpurun_out/bench_n${N}_eager.log 2>&1
echo "rc=$?" >> gpurun_out/bench_n${N}_eager.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 --no-roofline > gpurun_out/bench_n${N}_graph.log 2>&1
echo "rc=$?" >> gpurun_out/bench_n${N}_graph.log
grep -E "bench rank|rc=|metric" gpurun_out/bench_n${N}_eager.log | cut -c1-300 | tail -12
grep -E "bench rank|rc=|metric" gpurun_out/bench_n${N}_graph.log | cut -c1-300 | tail -12
