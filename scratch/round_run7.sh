#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_strip_mode.py tests/test_real_pair.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python scratch/prof.py c3 2>&1 | head -7
for d in 4 8; do timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --depth $d 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth', d['config']['in_flight_slots'], round(d['value']), [(k['name'], round(k['ms_per_step']*1e3,1)) for k in d['kernels'][:5]])"; done
