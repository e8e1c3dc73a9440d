#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_batch_gpu.py -m gpu -x -q 2>&1 | tail -n 5
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/full_tests.log
for v in 1 2; do
  timeout 300 python bench.py --workload c3 --variant $v --no-cpu-baseline --also "" --steps 5 > gpurun_out/c3_v$v.json 2>/dev/null
  python scripts/show_bench.py gpurun_out/c3_v$v.json 2>/dev/null | head -n 1
done
