#!/bin/bash
# independent-stream scaling point on N GPUs (gpurun --gpus N): bench.py under torchrun, as the driver runs it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N --steps 50 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
print("N=$N value %.4g ms/step %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
