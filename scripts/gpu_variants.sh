#!/bin/bash
# bench every tuning variant under variants/ (C2, device-resident + isolated launch)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
W=${1:-c2}
for lib in default $(ls variants/*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset STITCHB200_LIB; else export STITCHB200_LIB=$PWD/$lib; fi
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/v.json 2> gpurun_out/v.err
  python - "$lib" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/v.json").read().strip().splitlines()[-1])
    print("%-40s value %.0f  isolated us %.1f  frac %.3f  timed %.3f" % (sys.argv[1], d["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d["roofline"]["timed_region"]["frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/v.err").read()[-300:])
PY
done
