#!/bin/bash
# round-2 evidence: ncu --set full of the frame kernels (single-frame launches, default structure)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_fs2$' -s 6 -c 2 -o gpurun_out/r02_fs2_c2 python scripts/frame_prof.py c2 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_fs2$' -s 6 -c 2 -o gpurun_out/r02_fs2_app6 python scripts/frame_prof.py app6 > gpurun_out/ncu_b.log 2>&1
for f in a b; do tail -n 2 gpurun_out/ncu_$f.log | cut -c1-300; done
