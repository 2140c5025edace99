#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(ROUNDS=1 timeout 600 bash tools/ab_bench.sh tools/ab/liblgx_old.so tools/ab/liblgx_t256.so tools/ab/liblgx_t256s4.so tools/ab/liblgx_t128s4.so tools/ab/liblgx_t512s8.so tools/ab/liblgx_t256s16.so cylinder-pose-estimation_b200/liblgx.so 2>&1) > $O/s6_ab.log
(timeout 200 python tools/ridge_phase_prof.py 128 16 2>&1 | tail -8) > $O/s6_roles.log
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:blur5_u8_kernel|morph_kernel|jl_union|jl_sums|jl_roots|emit_kernel|jl_rank_assign' -c 8 -f -o $O/prof_s6 python tools/ridge_ws_prof.py 64 16 1 > $O/s6_ncu_full.log 2>&1
cat $O/s6_ab.log $O/s6_roles.log; tail -2 $O/s6_ncu_full.log
