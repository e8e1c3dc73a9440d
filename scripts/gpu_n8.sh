#!/bin/bash
# 8-GPU checks: the bench's scaling line (C2) and the strip (latency) mode
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -2 gpurun_out/r02_bench_n$N.err | cut -c1-300
python scripts/show_bench.py gpurun_out/r02_bench_n$N.json | head -2
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "copy ceiling", d["e2e"]["copy_only_ceiling"])
PY
bash scripts/gpu_strip.sh $N
