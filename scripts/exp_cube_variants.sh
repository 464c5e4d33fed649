#!/bin/bash
# development: scripts/exp_cube.py for every build variant pynmap_b200/libpnm_t*.so, then one ncu capture of the default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r02_exp_cube_variants.log
for lib in pynmap_b200/libpnm_t*.so; do
  echo "=== $lib" >> gpurun_out/r02_exp_cube_variants.log
  PNM_B200_LIB=$PWD/$lib timeout 300 python scripts/exp_cube.py 1e8 >> gpurun_out/r02_exp_cube_variants.log 2>&1
done
if [ "$1" = "ncu" ]; then bash scripts/ncu_cube.sh; fi
