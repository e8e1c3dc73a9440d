#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_fs2$' -s 4 -c 1 -o gpurun_out/r02_mbf_c3 python scripts/frame_prof.py c3 > gpurun_out/ncu_e.log 2>&1
tail -n 2 gpurun_out/ncu_e.log | cut -c1-300
