#!/bin/bash
# ncu --set full of the frame kernels (single-frame launches, default structure): the captures summarised under profiles/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:^k_fs2$' -s 6 -c 1 -f -o gpurun_out/r02_fs2_c2 python scripts/frame_prof.py c2 > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:^k_fs2$' -s 6 -c 1 -f -o gpurun_out/r02_fs2_app6 python scripts/frame_prof.py app6 > gpurun_out/ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_fs2$|k_mb_(pyr|band)' -s 36 -c 12 -f -o gpurun_out/r02_mb_c3 python scripts/frame_prof.py c3 > gpurun_out/ncu_c.log 2>&1
for f in a b c; do tail -n 1 gpurun_out/ncu_$f.log | cut -c1-200; done
