#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mb -s 24 -c 4 -o gpurun_out/r02_mb_v1 python scripts/mb_prof.py c3 > gpurun_out/ncu_mb.log 2>&1
tail -3 gpurun_out/ncu_mb.log
