#!/bin/bash
# N-GPU sweep: does leaving SMs to NCCL (fewer NCCL CTAs, a smaller wgrad stream-K partition) reduce the exposed exchange?
N=${1:-4}
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus $N --steps 20 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/n${N}_$name.log 2>&1
  echo -n "$name: "; grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/n${N}_$name.log; }
echo -n "single: "; timeout 200 python bench.py --steps 20 --warmup 3 --no-roofline --no-cpu-baseline 2>/dev/null | grep -o "\"ms_per_step\": [0-9.]*"
run default A=1
run ctas16 NCCL_MAX_CTAS=16
run ctas16_wp66 NCCL_MAX_CTAS=16 B200GAN_WGRAD_PAIRS=66
run ctas8_wp70 NCCL_MAX_CTAS=8 B200GAN_WGRAD_PAIRS=70
run nooverlap B200GAN_OVERLAP_UPDATE=0 B200GAN_BUCKET_MB=1000
run default2 A=1
