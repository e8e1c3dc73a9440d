#!/bin/bash
cd /root/repo
SB_FTS_NO_ROTATE=1 timeout 300 python scratch/fv_run.py c2 norotate 2>&1 | tail -1
timeout 300 python scratch/fv_run.py c2 rotate 2>&1 | tail -1
SB_FTS_NO_ROTATE=1 timeout 300 python scratch/fv_run.py c2 norotate 2>&1 | tail -1
timeout 300 python scratch/fv_run.py c2 rotate 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b_c2.json 2> gpurun_out/b_c2.err; python -c "
import json; d=json.loads(open('gpurun_out/b_c2.json').read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['frac'], d['roofline']['timed_region'])"
SB_FTS_NO_ROTATE=1 timeout 600 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('norotate', round(d['value']), d['roofline']['frac'])"
