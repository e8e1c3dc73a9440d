#!/bin/bash
cd /root/repo
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b_c2.json 2> gpurun_out/b_c2.err; tail -c 200 gpurun_out/b_c2.json
timeout 600 python bench.py --workload c3 --steps 20 --warmup 3 > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err; tail -c 200 gpurun_out/b_c3.json
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/b_c4.json 2> gpurun_out/b_c4.err; tail -c 200 gpurun_out/b_c4.json
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_mb_' -s 10 -c 5 -f -o gpurun_out/r01_mb_v5 python scripts/ncu_frame.py c3 4 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 > gpurun_out/launches_c3.out 2>&1
