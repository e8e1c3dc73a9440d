#!/bin/bash
cd /root/repo
for v in old h32w16 h16w8 h32w8 h16w16 h32w16p2 h64w16; do
  STITCHB200_LIB=/root/repo/scratch/variants/libstitchb200_$v.so timeout 300 python scratch/fv_run.py c2 $v 2>&1 | tail -2
done
