#!/bin/bash
# quick GPU visit: parity tests + short bench (no CPU leg)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
r=d['roofline']
print('value %.4g  ms/step %.4f dev_ms %.4f  e2e %.4g' % (d['value'], d['ms_per_step'], d['device_ms_per_step'], d['e2e']['value']))
print('decode ms %.4f  classify ms %.4f frac %.3f  stage frac %.3f  share %s' % (r['ms_per_launch'], r['sync_search']['ms_per_launch'], r['sync_search']['frac'], r['descramble_deinterleave_stage']['frac'], r['step_share']))
PY
