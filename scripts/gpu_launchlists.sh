#!/bin/bash
# launch lists of the bench command, one per workload (ncu --metrics gpu__time_duration.sum): the kernels' SHARES of a step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for w in c2 c3 app6; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_$w.csv python bench.py --workload $w --also "" --steps 1 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/ncu_l_$w.log 2>&1
  wc -l gpurun_out/r02_launches_$w.csv
done
