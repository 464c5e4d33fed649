#!/bin/bash
# one full ncu capture of k_project_cube (float64 accumulators) on the bench workload + the launch list of the script
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_project_cube --launch-skip 12 --launch-count 1 \
  -f -o gpurun_out/r02_cube python scripts/exp_cube.py 1e8 > gpurun_out/r02_cube_ncu.log 2>&1
echo "exit $?" >> gpurun_out/r02_cube_ncu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_cube_launches.csv \
  python scripts/exp_cube.py 1e8 > gpurun_out/r02_cube_launches.log 2>&1
