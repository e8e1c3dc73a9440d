#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/full_tests.log
timeout 900 python bench.py > gpurun_out/r02_final_b_n1.json 2> gpurun_out/r02_final_b_n1.err; echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 300 python bench.py --workload c5-strip --steps 40 2>/dev/null | tail -n 1 > gpurun_out/r02_final_strip_n1.json; cut -c1-200 gpurun_out/r02_final_strip_n1.json
