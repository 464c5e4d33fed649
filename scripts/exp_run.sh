#!/bin/bash
cd "$(dirname "$0")/.."
for t in 0 3 4 5 6; do PNM_TMA_SLOTS=$t python scripts/exp_knobs.py 1e8; done
