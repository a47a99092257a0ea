#!/bin/bash
# quick GPU visit: parity tests + probe both Viterbi forms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_probe.py 1000000 2>&1 | tail -14
