#!/bin/bash
# development sweep (run under gpurun)
cd "$(dirname "$0")/.."
L=$PWD/pynmap_b200
for mb in 2 3 4; do
  PNM_B200_LIB=$L/libpnm_mb$mb.so python scripts/exp_knobs.py 1e8 --sorted
done
PNM_B200_LIB=$L/libpnm_mb2.so PNM_CTAS_PER_SM=4 python scripts/exp_knobs.py 1e8
PNM_B200_LIB=$L/libpnm_mb3.so PNM_CTAS_PER_SM=6 python scripts/exp_knobs.py 1e8
for r in 1 8 64 128; do
  PNM_MAX_REPLICAS=$r PNM_REPLICA_MIB=128 python scripts/exp_knobs.py 1e8
done
PNM_REP_SHIFT=0 python scripts/exp_knobs.py 1e8 --sorted
PNM_REP_SHIFT=0 PNM_MAX_REPLICAS=128 PNM_REPLICA_MIB=128 python scripts/exp_knobs.py 1e8
PNM_REP_SHIFT=3 python scripts/exp_knobs.py 1e8
