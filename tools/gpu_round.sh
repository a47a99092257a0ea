#!/bin/bash
# one GPU visit: parity tests, bench, ncu launch list, ncu full capture of the dominant kernel
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
VIT=${VIT:-1}
timeout 600 python bench.py --viterbi $VIT > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 600 gpurun_out/bench_ref.json
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --viterbi $VIT > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_decode_lane|k_classify_tile' -s 2 -c 2 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --viterbi $VIT > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
fi
