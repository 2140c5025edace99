#!/bin/bash
# end-of-round evidence (round 2): full GPU test-suite, smoke, both bench arms, launch list of the bench command,
# ncu --set full of the dominant kernels (each ncu command runs right after the same command has exited 0 without ncu)
mkdir -p gpurun_out
O=gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/g_pytest.log
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -4) > $O/g_smoke.log
timeout 600 python bench.py > $O/g_bench.json 2> $O/g_bench.err
timeout 400 python bench.py --impl reference > $O/g_bench_ref.json 2> $O/g_bench_ref.err
K='regex:ridge|sauvola|blur5|morph|jl_|emit|fill_holes|pack_bits|unpack_bits|bgr2gray|undistort'
python bench.py --steps 2 --warmup 1 --no-cpu --check 0 > $O/g_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file $O/g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --check 0 > $O/g_ncu.log 2>&1
python tools/ridge_ws_prof.py 74 16 1 > $O/g_plain2.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:ridge_ws_kernel|sauvola_kernel' -c 2 -f -o $O/prof_g python tools/ridge_ws_prof.py 74 16 1 > $O/g_ncu_full.log 2>&1
python tools/joints_prof.py 32 32 > $O/g_plain3.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:jl_local' -s 3 -c 1 -f -o $O/prof_g_jl python tools/joints_prof.py 32 32 > $O/g_ncu_jl.log 2>&1
cat $O/g_pytest.log $O/g_smoke.log $O/g_bench.json $O/g_bench_ref.json; tail -n 2 $O/g_ncu_full.log $O/g_ncu_jl.log
