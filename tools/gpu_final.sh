#!/bin/bash
# last visit of a round: parity tests, full bench line, ncu launch list, one full capture of the decode kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 100 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:'k_decode_lane|k_classify_tile' -s 2 -c 2 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/ | head -20
