#!/bin/bash
# A/B: run the short device-resident bench with each prebuilt library variant in tools/variants/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp osmo-tetra_b200/libtetra_b200.so /tmp/keep.so
for v in tools/variants/*.so; do
  cp $v osmo-tetra_b200/libtetra_b200.so
  echo "== $v"
  timeout 300 python bench.py --no-cpu --no-e2e --steps 50 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.4g dev_ms %.4f decode %.4f classify %.4f' % (d['value'], d['device_ms_per_step'], r['ms_per_launch'], r['sync_search']['ms_per_launch']))"
done
cp /tmp/keep.so osmo-tetra_b200/libtetra_b200.so
