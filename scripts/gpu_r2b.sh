#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_batch_gpu.py tests/test_strip_mode.py -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b.json 2> gpurun_out/r2b.err
tail -3 gpurun_out/r2b.err
python scripts/show_bench.py gpurun_out/r2b.json
