#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest tests/test_undistort.py -m gpu -x -q 2>&1 | tail -8) > $O/s8_pytest.log
(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --check 2 2>$O/s8_bench.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value'])); print(json.dumps(d['undistort']))") > $O/s8_bench.log 2>&1
(tools/microbench/fp64_mix; tools/microbench/fp64_rate) > $O/s8_microbench.log 2>&1
cat $O/s8_pytest.log $O/s8_bench.log; tail -3 $O/s8_bench.err; cat $O/s8_microbench.log
