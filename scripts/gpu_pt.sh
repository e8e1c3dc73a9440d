#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for lib in default $(ls variants/*.so); do
  if [ "$lib" = default ]; then unset STITCHB200_LIB; else export STITCHB200_LIB=$PWD/$lib; fi
  echo "== $lib"
  timeout 300 python bench.py --workload c3 --no-cpu-baseline --also "" --steps 5 > gpurun_out/pt.json 2>/dev/null
  python scripts/show_bench.py gpurun_out/pt.json 2>/dev/null | grep -E "C3|pyr_down"
done
