#!/bin/bash
# round-2 microbenchmark batch: writes gpurun_out/r02_micro_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem --format=csv > gpurun_out/r02_micro_env.log 2>&1
for b in red_limits_bench red_pair_bench tma_mix_bench tma_red_bench tma_load_bench; do
  timeout 120 scripts/$b > gpurun_out/r02_micro_$b.log 2>&1
  echo "exit $?" >> gpurun_out/r02_micro_$b.log
done
