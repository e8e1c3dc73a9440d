#!/bin/bash
# N-GPU checks: the bench's scaling line (C2) and the latency (strip) mode through bench.py --workload c5-strip
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_final_b_n$N.json 2> gpurun_out/r02_final_b_n$N.err
tail -n 2 gpurun_out/r02_final_b_n$N.err | cut -c1-300
python scripts/show_bench.py gpurun_out/r02_final_b_n$N.json | head -n 2
p=29520
for halo in recompute peer exchange; do
  p=$((p+1))
  timeout 400 $TR --master-port $p bench.py --gpus $N --workload c5-strip --halo $halo --steps 40 --warmup 5 > gpurun_out/r02_final_strip_n${N}_$halo.json 2> gpurun_out/r02_final_strip_n${N}_$halo.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_final_strip_n${N}_$halo.json").read().strip().splitlines()[-1])
    print("$halo", d["n_gpus"], "latency ms", d["latency_ms"], "bit exact", d["bit_exact_vs_unsplit"])
except Exception as e:
    print("$halo failed", e)
PY
done
