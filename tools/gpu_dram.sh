#!/bin/bash
# DRAM traffic + duration of the kernels matching $KERN, second launch; then the quick bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"${KERN:-k_decode_lane}" -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep -E "dram__|gpu__time|lts__|k_" 
timeout 600 python bench.py --no-cpu --no-e2e --steps 50 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.4g dev_ms %.4f decode %.4f search %.4f' % (d['value'], d['device_ms_per_step'], r['ms_per_launch'], r['sync_search']['ms_per_launch']))"
