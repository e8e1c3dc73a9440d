#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-2}
p=29540
for h in ${HALOS:-recompute peer exchange}; do
  p=$((p+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $p bench.py --gpus $N --workload c5-strip --halo $h --steps 40 --warmup 5 2>gpurun_out/strip_$h.err | tail -n 1 > gpurun_out/r02_final_strip_n${N}_$h.json
  python -c "import json;d=json.loads(open('gpurun_out/r02_final_strip_n${N}_$h.json').read());print('$h',d['n_gpus'],d['latency_ms'],d['bit_exact_vs_unsplit'])" || tail -n 5 gpurun_out/strip_$h.err
done
