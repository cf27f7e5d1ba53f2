#!/bin/bash
# 2-GPU sweep of the exchange knobs; logs under gpurun_out/
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus 2 --steps 20 --warmup 3 --no-roofline > gpurun_out/n2_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
for l in open("gpurun_out/n2_%s.log" % sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], "ms/step %.3f  value %.0f  e2e %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
}
timeout 200 python bench.py --steps 20 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/n2_single.log 2>&1
python -c "
import json
for l in open('gpurun_out/n2_single.log'):
    if l.startswith('{'): d=json.loads(l); print('single ms/step %.3f' % d['ms_per_step'])"
run default A=1
run prio B200GAN_SIDE_PRIORITY=1
run prio_ctas16 B200GAN_SIDE_PRIORITY=1 NCCL_MAX_CTAS=16
run prio_ctas32 B200GAN_SIDE_PRIORITY=1 NCCL_MAX_CTAS=32
run prio_onebucket B200GAN_SIDE_PRIORITY=1 B200GAN_BUCKET_MB=1000
run default2 A=1
