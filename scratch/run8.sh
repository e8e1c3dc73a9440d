#!/bin/bash
cd /root/repo
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python scratch/prof.py c3 2>&1 | head -7
