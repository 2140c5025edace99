#!/bin/bash
# final check of the tree as committed: full GPU test-suite, smoke, both bench arms, launch list
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/f_pytest.log
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3) > $O/f_smoke.log
timeout 400 python bench.py > $O/f_bench.json 2> $O/f_bench.err
timeout 400 python bench.py --impl reference > $O/f_bench_ref.json 2> $O/f_bench_ref.err
K='regex:ridge|sauvola|blur5|morph|jl_|emit|fill_holes|pack_bits|bgr2gray|undistort'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file $O/f_launches.csv python bench.py --steps 1 --warmup 1 --batch 128 --chunk 64 --no-cpu --check 0 > $O/f_ncu.log 2>&1
cat $O/f_pytest.log $O/f_smoke.log; cut -c1-330 $O/f_bench.json; tail -2 $O/f_bench.err
