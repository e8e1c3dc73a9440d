#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compositor or blocks" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_batch_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload c3 --also "" > gpurun_out/mb_c3.json 2> gpurun_out/mb_c3.err
tail -2 gpurun_out/mb_c3.err
python scripts/show_bench.py gpurun_out/mb_c3.json
