#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_batch_gpu.py tests/test_real_pair.py tests/test_real_pair_full.py tests/test_strip_mode.py -m gpu -x -q 2>&1 | tail -n 8
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -n 5
timeout 300 python scripts/mb_prof.py c3 12,18 2>&1 | tail -n 8
