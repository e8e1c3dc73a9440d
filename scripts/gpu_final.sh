#!/bin/bash
# final evidence of the round: full GPU suite, the two bench arms, launch list of the bench command, ncu --set full of the frame kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/full_tests.log
timeout 900 python bench.py > gpurun_out/r02_final_b_n1.json 2> gpurun_out/r02_final_b_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r02_final_b_ref.json 2> gpurun_out/r02_final_b_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_fs2$' -s 6 -c 1 -o gpurun_out/r02_fs2_c2 -f python scripts/frame_prof.py c2 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_fs2$' -s 6 -c 1 -o gpurun_out/r02_fs2_app6 -f python scripts/frame_prof.py app6 > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_fs2$|k_mb_(pyr|band)' -s 36 -c 12 -o gpurun_out/r02_mb_c3 -f python scripts/frame_prof.py c3 > gpurun_out/ncu_c.log 2>&1
for f in a b c d; do tail -n 1 gpurun_out/ncu_$f.log | cut -c1-200; done
