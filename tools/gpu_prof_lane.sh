#!/bin/bash
# ncu --set full of the split decode pass (k_lane_prepare, k_lane_trellis) on config 4 and config 2: tag = $1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02b}
N=${PROF_SLOTS:-2097152}
for shape in config4 config2; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_lane_prepare|k_lane_trellis|k_lane_finish' -s 2 -c 2 -f \
    -o gpurun_out/prof_${TAG}_$shape python tools/prof_run.py $shape $N 2 > gpurun_out/ncu_${TAG}_$shape.log 2>&1
  grep PROF gpurun_out/ncu_${TAG}_$shape.log
  python tools/prof_run.py $shape $N 4 | grep PROF
done
ls -la gpurun_out/*.ncu-rep
