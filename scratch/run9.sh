#!/bin/bash
cd /root/repo
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python scratch/fv_run.py c2 hardened 2>&1 | tail -1
