#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
for v in 0 1; do
  echo "== SB_PDL=$v"
  SB_PDL=$v timeout 300 python bench.py --no-cpu-baseline --also app6 --steps 10 > gpurun_out/q.json 2>/dev/null; python scripts/show_bench.py gpurun_out/q.json 2>/dev/null | grep -v "^ "
done
