#!/bin/bash
# round 2, first GPU call: parity of the new feather kernel + a first bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compositor" 2>&1 | tail -15 > gpurun_out/r2a_tests.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "full_size" 2>&1 | tail -15 >> gpurun_out/r2a_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_c2.json 2> gpurun_out/r2a_c2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variant 5 > gpurun_out/r2a_c2_v1.json 2> gpurun_out/r2a_c2_v1.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload app6 > gpurun_out/r2a_app6.json 2> gpurun_out/r2a_app6.err
cat gpurun_out/r2a_tests.log
python - <<'PY'
import json
for f in ("r2a_c2","r2a_c2_v1","r2a_app6"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f frac %.3f us %.1f timed %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["roofline"]["timed_region"]["frac"]))
    except Exception as e:
        print(f, "failed", e)
PY
