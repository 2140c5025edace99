#!/bin/bash
# end-of-round evidence: full GPU test-suite, smoke, both bench arms, launch list, ncu --set full of the main kernels
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/g_pytest.log
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3) > $O/g_smoke.log
timeout 400 python bench.py > $O/g_bench.json 2> $O/g_bench.err
timeout 400 python bench.py --impl reference > $O/g_bench_ref.json 2> $O/g_bench_ref.err
K='regex:ridge|sauvola|blur5|morph|jl_|emit|fill_holes|pack_bits|bgr2gray|undistort'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file $O/g_launches.csv python bench.py --steps 1 --warmup 1 --batch 128 --chunk 64 --no-cpu --check 0 > $O/g_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:ridge_ws_kernel|sauvola_kernel|blur5_u8_kernel' -c 3 -f -o $O/prof_g python tools/ridge_ws_prof.py 74 16 1 > $O/g_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:undistort_kernel' -c 1 -f -o $O/prof_g_und python tools/undistort_prof.py 64 > $O/g_ncu_und.log 2>&1
(timeout 200 python tools/config_sweep.py 2>&1 | tail -12) > $O/g_sweep.log
cat $O/g_pytest.log $O/g_smoke.log $O/g_bench.json; tail -2 $O/g_ncu_full.log $O/g_ncu_und.log; cat $O/g_sweep.log
