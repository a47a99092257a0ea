#!/bin/bash
# A/B on one box: short device-resident bench for every "name|library|ENV=.. ENV=.." line of tools/ab_specs.txt
# (library: path of a prebuilt variant, or - for the in-tree build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp osmo-tetra_b200/libtetra_b200.so /tmp/keep.so
while IFS='|' read -r name lib envs; do
  [ -z "$name" ] && continue
  case "$name" in \#*) continue;; esac
  if [ "$lib" != "-" ]; then cp "$lib" osmo-tetra_b200/libtetra_b200.so; else cp /tmp/keep.so osmo-tetra_b200/libtetra_b200.so; fi
  out=$(env $envs timeout 300 python bench.py --no-cpu --no-e2e --steps ${AB_STEPS:-50} ${AB_ARGS:-} 2>gpurun_out/ab_err.log | tail -1)
  echo "$out" | python -c "
import json,sys
name=sys.argv[1]
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
    print('%-28s value %.4g  ms/step %.4f dev_ms %.4f decode %.4f search %.4f share %s' % (name, d['value'], d['ms_per_step'], d['device_ms_per_step'], r['ms_per_launch'], r['sync_search']['ms_per_launch'], {k: round(v,3) for k,v in r['step_share'].items()}))
except Exception as e:
    print(name, 'FAILED', e); print(open('gpurun_out/ab_err.log').read()[-800:])
" "$name"
done < tools/ab_specs.txt
cp /tmp/keep.so osmo-tetra_b200/libtetra_b200.so
