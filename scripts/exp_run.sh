#!/bin/bash
cd "$(dirname "$0")/.."
L=$PWD/pynmap_b200
python scripts/exp_knobs.py 1e8
for t in 0 2 3; do PNM_B200_LIB=$L/libpnm_mb3.so PNM_TMA_SLOTS=$t python scripts/exp_knobs.py 1e8; done
PNM_B200_LIB=$L/libpnm_mb3.so PNM_TMA_SLOTS=3 PNM_CTAS_PER_SM=48 python scripts/exp_knobs.py 1e8
