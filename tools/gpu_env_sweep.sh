#!/bin/bash
# sweep an environment variable over values on one box: VAR=name VALS="a b c" bash tools/gpu_env_sweep.sh
cd "$(dirname "$0")/.."
for v in $VALS; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 300 python bench.py --no-cpu --no-e2e --steps 50 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.4g dev_ms %.4f decode %.4f search %.4f' % (d['value'], d['device_ms_per_step'], r['ms_per_launch'], r['sync_search']['ms_per_launch']))"
done
