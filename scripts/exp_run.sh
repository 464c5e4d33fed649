#!/bin/bash
cd "$(dirname "$0")/.."
python scripts/exp_knobs.py 1e8 --sorted
