#!/bin/bash
# development: 8-GPU bench lines for a few settings of the reduce (one JSON line each -> gpurun_out/r02_n8_variants.jsonl)
N=${1:-8}
mkdir -p gpurun_out
: > gpurun_out/r02_n${N}_variants.jsonl
port=29600
run() {
  port=$((port + 1))
  echo "# $*" >> gpurun_out/r02_n${N}_variants.jsonl
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --no-e2e --no-extras --no-cpu-baseline "$@" 2> gpurun_out/n${N}_last.err | grep '^{' >> gpurun_out/r02_n${N}_variants.jsonl
}
run --steps 60 --warmup 5 --no-parity --nccl-max-ctas 4
run --steps 60 --warmup 5 --no-parity --nccl-max-ctas 8
run --steps 60 --warmup 5 --no-parity --nccl-max-ctas 16
run --steps 60 --warmup 5 --no-parity --nccl-max-ctas 2
run --steps 60 --warmup 5 --no-parity --nccl-max-ctas 8 --reserve-sms 0
run --steps 20 --warmup 5 --nccl-max-ctas 8
python - <<PY
import json
for l in open("gpurun_out/r02_n${N}_variants.jsonl"):
    if l.startswith("#"): print(l.strip()); continue
    j = json.loads(l)
    print("  ms/step %.4f  kernel_ms %.4f  value %.4g  host_enqueue %.3f  parity %s" % (j["ms_per_step"], j["roofline"]["kernel_ms"], j["value"], j["host_enqueue_ms_per_step"], (j.get("parity") or {}).get("equals_1rank")))
PY
tail -3 gpurun_out/n${N}_last.err
