#!/bin/bash
# round-2 evidence on one GPU: -m gpu tests, smoke, the default bench line, then (after the un-profiled run exited 0) the
# ncu launch list and one full capture of k_project_cube from the same bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_gputests_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "reference rc=$?"
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-parity"
$CMD > gpurun_out/r02_bench_same_command.json 2>/dev/null && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 60 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_project_cube --launch-skip 4 --launch-count 1 \
  -f -o gpurun_out/r02_cube_final $CMD > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"
