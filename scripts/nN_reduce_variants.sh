#!/bin/bash
# development: N-GPU bench lines for the three flavours of the reduction over ranks -> gpurun_out/r02_n${N}_reduce.jsonl
N=${1:-2}
mkdir -p gpurun_out
out=gpurun_out/r02_n${N}_reduce.jsonl
: > $out
port=29800
run() {
  port=$((port + 1))
  echo "# $*" >> $out
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --no-e2e --no-extras --no-cpu-baseline "$@" 2> gpurun_out/n${N}_last.err | grep '^{' >> $out || tail -20 gpurun_out/n${N}_last.err
}
if [ "$2" = "quick" ]; then
run --steps 20 --warmup 5 --reduce peer
else
run --steps 20 --warmup 5 --reduce peer
run --steps 20 --warmup 5 --no-parity --reduce peer
run --steps 60 --warmup 5 --no-parity --reduce peer
fi
python - <<PY
import json
for l in open("$out"):
    if l.startswith("#"): print(l.strip()); continue
    j = json.loads(l)
    print("  ms/step %.4f  kernel_ms %.4f  value %.4g  host_enqueue %.3f  parity %s" % (j["ms_per_step"], j["roofline"]["kernel_ms"], j["value"], j["host_enqueue_ms_per_step"], (j.get("parity") or {}).get("equals_1rank")))
PY
