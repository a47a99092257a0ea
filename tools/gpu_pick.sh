#!/bin/bash
# A/B on one box, then the parity tests with the fastest variant if it is not the baseline (tools/variants/v0.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp osmo-tetra_b200/libtetra_b200.so /tmp/keep.so
: > gpurun_out/pick.txt
for v in tools/variants/*.so; do
  cp $v osmo-tetra_b200/libtetra_b200.so
  timeout 60 python bench.py --no-cpu --no-e2e --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v %.4f %.4g' % (r['ms_per_launch'], d['value']))" | tee -a gpurun_out/pick.txt
done
best=$(sort -k2 -n gpurun_out/pick.txt | head -1 | cut -d' ' -f1)
echo "best $best" | tee -a gpurun_out/pick.txt
if [ "$best" != "tools/variants/v0.so" ]; then
  cp $best osmo-tetra_b200/libtetra_b200.so
  timeout 70 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee -a gpurun_out/pick.txt
fi
cp /tmp/keep.so osmo-tetra_b200/libtetra_b200.so
