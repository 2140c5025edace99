#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q -k "blur5 or frontend_cylinder or golden" 2>&1 | tail -3) > $O/j_pytest.log
(ROUNDS=2 timeout 500 bash tools/ab_bench.sh tools/ab/liblgx_old.so cylinder-pose-estimation_b200/liblgx.so tools/ab/liblgx_pf8.so tools/ab/liblgx_pf24.so tools/ab/liblgx_pf64.so 2>&1) > $O/j_ab.log
cat $O/j_pytest.log $O/j_ab.log
