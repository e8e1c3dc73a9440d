#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_batch_gpu.py tests/test_real_pair.py tests/test_strip_mode.py -m gpu -x -q 2>&1 | tail -n 3
timeout 300 python bench.py --workload c3 --no-cpu-baseline --also "" --steps 5 > gpurun_out/pt.json 2>/dev/null
python scripts/show_bench.py gpurun_out/pt.json 2>/dev/null | head -n 6
