#!/bin/bash
# one point of the 1/2/4/8 scaling series of the default workload (C2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_final_b_n$N.json 2> gpurun_out/r02_final_b_n$N.err
python scripts/show_bench.py gpurun_out/r02_final_b_n$N.json | head -n 1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c5-strip --steps 40 --warmup 5 2>/dev/null | tail -n 1 > gpurun_out/r02_final_strip_n${N}_recompute.json
cut -c1-220 gpurun_out/r02_final_strip_n${N}_recompute.json
