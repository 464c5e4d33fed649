#!/bin/bash
cd "$(dirname "$0")/.."
for c in 8 16 32 64 128; do PNM_CTAS_PER_SM=$c python scripts/exp_knobs.py 1e8 --sorted; done
for c in 16 32 64; do PNM_PIPE=1 PNM_CTAS_PER_SM=$c python scripts/exp_knobs.py 1e8; done
for c in 24 48 96; do PNM_PIPE=4 PNM_CTAS_PER_SM=$c python scripts/exp_knobs.py 1e8; done
