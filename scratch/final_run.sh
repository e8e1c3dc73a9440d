#!/bin/bash
cd /root/repo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b_c2.json 2> gpurun_out/b_c2.err
timeout 600 python bench.py --workload c3 --steps 20 --warmup 3 > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/b_ref_c2.json 2> gpurun_out/b_ref.err
python - <<'P'
import json
for f in ('gpurun_out/b_c2.json','gpurun_out/b_c3.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'depth', d['config']['in_flight_slots'], 'frac', round(d['roofline']['frac'],3), 'timed', round(d['roofline']['timed_region']['frac'],3), 'launches', d['gpu_launches'], d['clocks'])
d=json.loads(open('gpurun_out/b_ref_c2.json').read().strip().splitlines()[-1]); print('ref', d['value'], d['cpu_baseline']['cores'])
P
