#!/bin/bash
# 8-GPU bench with the NCCL reduce capped to a few CTAs (bench.py --nccl-max-ctas); one JSON line per setting.
mkdir -p gpurun_out
for c in 0 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29510 + c)) \
    bench.py --gpus 8 --steps 60 --warmup 8 --no-e2e --no-extras --no-cpu-baseline --nccl-max-ctas $c \
    2> gpurun_out/sweep_$c.err | grep '^{' | tee gpurun_out/sweep_$c.json | python -c "
import json,sys
for l in sys.stdin:
    j=json.loads(l); print('max_ctas=$c', 'ms/step', j['ms_per_step'], 'value %.4g' % j['value'], 'kernel_ms', j.get('kernel_ms'))"
done
