#!/bin/bash
# one GPU call: variant check, both bench arms on C2, C3 bench, launch lists
cd /root/repo
timeout 300 python scratch/fv_run.py c2 v14 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b_c2.json 2> gpurun_out/b_c2.err; tail -c 600 gpurun_out/b_c2.json
timeout 600 python bench.py --workload c3 --steps 20 --warmup 3 > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err; tail -c 300 gpurun_out/b_c3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_c2.out 2>&1
