#!/bin/bash
# round-2 captures of the final tree: launch list of a short bench run + ncu --set full of the hot kernels
# (k_classify_tile, k_sb1_lane, k_lane_prepare, k_lane_trellis) on config 4 and config 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${PROF_SLOTS:-2097152}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv \
  python bench.py --steps 2 --warmup 1 --bursts 8000000 --no-e2e --no-cpu --no-config5 --no-configs --no-parity > gpurun_out/r02b_launches_bench.log 2>&1
tail -c 300 gpurun_out/r02b_launches_bench.log
for shape in config4 config2; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_lane_prepare|k_lane_trellis|k_classify_tile|k_sb1_lane' -s 4 -c 4 -f \
    -o gpurun_out/prof_r02b_$shape python tools/prof_run.py $shape $N 2 > gpurun_out/ncu_r02b_$shape.log 2>&1
  grep PROF gpurun_out/ncu_r02b_$shape.log
  python tools/prof_run.py $shape $N 4 | grep PROF | sed 's/^PROF/LIVE/'
done
ls -la gpurun_out/*.ncu-rep
