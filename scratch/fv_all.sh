#!/bin/bash
cd /root/repo
STITCHB200_LIB=/root/repo/scratch/variants/libstitchb200_h32w16.so timeout 300 python scratch/fv_run.py c2 v12 2>&1 | tail -1
timeout 300 python scratch/fv_run.py c2 v13 2>&1 | tail -1
SB_FTS_NO_SORT=1 timeout 300 python scratch/fv_run.py c2 v13-nosort 2>&1 | tail -1
timeout 300 python scratch/fv_run.py c2 v13 2>&1 | tail -1
