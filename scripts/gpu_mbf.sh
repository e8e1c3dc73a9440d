#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python scripts/mb_prof.py c3 12,17 2>&1 | tail -n 8
timeout 300 python bench.py --no-cpu-baseline --also app6 --steps 10 > gpurun_out/q.json 2>/dev/null; python scripts/show_bench.py gpurun_out/q.json 2>/dev/null | grep -v "^ "
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_batch_gpu.py tests/test_round2_features.py -m gpu -x -q 2>&1 | tail -n 3
