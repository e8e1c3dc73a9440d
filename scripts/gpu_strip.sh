#!/bin/bash
# strip (latency) mode across N GPUs: halo exchange by NCCL send/recv, by peer-memory writes, and halo recompute
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-2}
for halo in peer recompute exchange; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      scripts/strip_run.py --rig c5 --frames 20 --halo $halo > gpurun_out/strip_c5_n${N}_$halo.json 2> gpurun_out/strip_c5_n${N}_$halo.err
  echo "$halo rc=$?"; tail -1 gpurun_out/strip_c5_n${N}_$halo.json; tail -2 gpurun_out/strip_c5_n${N}_$halo.err | cut -c1-300
done
