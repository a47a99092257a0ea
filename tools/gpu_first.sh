#!/bin/bash
# first GPU contact: smoke (parity vs oracle), then a timing probe
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python tools/quick_probe.py 100000 2>&1 | tail -12
python tools/quick_probe.py 1000000 2>&1 | tail -12
