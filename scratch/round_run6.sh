#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_strip_mode.py tests/test_real_pair.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scratch/prof.py c3 2>&1 | head -7
