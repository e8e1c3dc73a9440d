#!/bin/bash
cd /root/repo
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python scratch/prof.py c3 2>&1 | head -7
timeout 400 python bench.py --workload app6 --steps 20 --warmup 3 > gpurun_out/b_app6.json 2> gpurun_out/b_app6.err; tail -2 gpurun_out/b_app6.err; python -c "
import json; d=json.loads(open('gpurun_out/b_app6.json').read().strip().splitlines()[-1]); print('app6', round(d['value']), 'e2e', round(d['e2e']['value']), 'vs_baseline', d['vs_baseline'], d['roofline']['kernel'], d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])"
