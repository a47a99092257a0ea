#!/bin/bash
# multi-GPU visit (gpurun --gpus N): NCCL sharding tests, independent-stream bench and the config-5 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_dist_nccl.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload config5 --total-bursts ${TOTAL:-100000000} --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_config5_n$N.json
python - <<PY
import json
for f in ("gpurun_out/bench_n$N.json", "gpurun_out/bench_config5_n$N.json"):
    try:
        d = json.load(open(f))
        print(f, "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], d.get("e2e", {}).get("value"), d.get("decode_only"))
    except Exception as e:
        print(f, "failed", e)
PY
