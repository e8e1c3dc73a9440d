#!/bin/bash
# quick loop: feather parity + C2 bench (+ optional ncu)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compositor" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "full_size" 2>&1 | tail -5
for w in c2 app6; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err
done
python - <<'PY'
import json
for f in ("q_c2","q_app6"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f frac %.3f us %.1f timed %.3f MB %.1f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["roofline"]["timed_region"]["frac"], d["roofline"]["algorithmic_bytes_per_launch"]/1e6))
    except Exception as e:
        print(f, "failed", e)
PY
if [ "$1" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fs2 -s 20 -c 1 -o gpurun_out/$2 python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/ncu_q.log 2>&1
fi
