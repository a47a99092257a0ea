#!/bin/bash
# round-2 captures: launch list of a short bench run + ncu --set full of the three hot kernels on config 4 and config 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${PROF_SLOTS:-2097152}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --bursts 8000000 --no-e2e --no-cpu --no-config5 --no-configs --no-parity > gpurun_out/r02_launches_bench.log 2>&1
tail -c 300 gpurun_out/r02_launches_bench.log
for shape in config4 config2; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_decode_lane|k_classify_tile|k_sb1_lane' -s 3 -c 3 -f \
    -o gpurun_out/prof_r02_$shape python tools/prof_run.py $shape $N 2 > gpurun_out/ncu_r02_$shape.log 2>&1
  grep PROF gpurun_out/ncu_r02_$shape.log
  python tools/prof_run.py $shape $N 4 | grep PROF
done
ls -la gpurun_out/*.ncu-rep
