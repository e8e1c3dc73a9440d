#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for m in 0 1; do
  echo "== SB_FS2_SCHED=$m"
  SB_FS2_SCHED=$m timeout 300 python bench.py --no-cpu-baseline --also app6 --steps 10 > gpurun_out/q.json 2>/dev/null; python scripts/show_bench.py gpurun_out/q.json 2>/dev/null | grep -v "^ "
done
SB_FS2_SCHED=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_batch_gpu.py -m gpu -x -q -k "feather or no or persistent or compositor" 2>&1 | tail -n 2
