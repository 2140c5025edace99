#!/bin/bash
# one GPU session: parity of the changed kernels, A/B against tools/ab/liblgx_old.so, e2e chunk sweep, launch list, ncu
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "stage1 or frontend or golden or strided or batch_equals or reference_named" 2>&1 | tail -5) > $O/s4_pytest.log
(timeout 300 bash tools/ab_bench.sh 2>&1) > $O/s4_ab.log
for ch in 16 32 64; do
  (timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --check 0 --e2e-chunk $ch 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('e2e_chunk', $ch, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'full', round(d['e2e']['with_u8_planes_back']['value']))") >> $O/s4_e2e.log 2>&1
done
K='regex:ridge|sauvola|blur5|morph|jl_|emit|fill_holes|pack_bits|bgr2gray'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file $O/s4_launches.csv python bench.py --steps 1 --warmup 1 --batch 128 --chunk 64 --no-cpu --check 0 > $O/s4_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:ridge_ws_kernel|sauvola_kernel' -c 2 -f -o $O/prof_s4 python tools/ridge_ws_prof.py 74 16 1 > $O/s4_ncu_full.log 2>&1
cat $O/s4_pytest.log $O/s4_ab.log $O/s4_e2e.log; tail -2 $O/s4_ncu_full.log
