#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scratch/fv_run.py c2 dp2a 2>&1 | tail -1
timeout 300 python scratch/prof.py c3 2>&1 | head -7
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_mb_|k_feather' -s 40 -c 14 -f -o gpurun_out/r01_mb_v5 python scripts/ncu_frame.py c3 6 2>&1 | tail -1
