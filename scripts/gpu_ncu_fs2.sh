#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fs2 -s 20 -c 2 -o gpurun_out/r02_fs2_v1 python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/ncu_fs2.log 2>&1
tail -3 gpurun_out/ncu_fs2.log
