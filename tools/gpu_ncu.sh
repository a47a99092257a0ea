#!/bin/bash
# ncu full capture of the kernels matching $KERN (regex), one launch each after warm-up
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KERN=${KERN:-k_classify_tile}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KERN" -s ${SKIP:-3} -c ${COUNT:-1} -f -o gpurun_out/prof_${TAG:-x} python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_${TAG:-x}.log 2>&1
tail -3 gpurun_out/ncu_${TAG:-x}.log
