#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "blur5 or frontend_small or frontend_cylinder or frontend_plane or golden or strided or config4" 2>&1 | tail -8) > $O/s7_pytest.log
(ROUNDS=2 timeout 600 bash tools/ab_bench.sh tools/ab/liblgx_old.so cylinder-pose-estimation_b200/liblgx.so tools/ab/liblgx_mb6.so tools/ab/liblgx_mb8.so 2>&1) > $O/s7_ab.log
for ch in 26 52; do
  (timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --check 0 --e2e-chunk $ch 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('e2e_chunk', $ch, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'full', round(d['e2e']['with_u8_planes_back']['value']))") >> $O/s7_e2e.log 2>&1
done
cat $O/s7_pytest.log $O/s7_ab.log $O/s7_e2e.log
