#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "undistort or batch_equals or frontend_small or golden" 2>&1 | tail -5) > $O/s9_pytest.log
for ch in 32 48 64 128; do
  (timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --check 0 --e2e-chunk $ch 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); u=d['undistort']; print('e2e_chunk', $ch, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'full', round(d['e2e']['with_u8_planes_back']['value']), 'undistort us/frame', round(u['ms_per_launch']/256*1e3,2))") >> $O/s9_e2e.log 2>&1
done
cat $O/s9_pytest.log $O/s9_e2e.log
