#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python bench.py --no-cpu-baseline --also app6 --steps 10 > gpurun_out/q.json 2>/dev/null; python scripts/show_bench.py gpurun_out/q.json 2>/dev/null | grep -v "^ "
STITCHB200_LIB=$PWD/variants/libstitchb200_trace.so timeout 300 python scripts/fs2_trace.py c2 stream4 -v 2>&1 | tail -n 50 > gpurun_out/trace.txt; head -n 2 gpurun_out/trace.txt
