#!/bin/bash
cd /root/repo
timeout 300 python scratch/fv_run.py c2 v14 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
