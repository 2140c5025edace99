#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "sauvola or stage1 or frontend_small or frontend_cylinder or golden or batch_equals or batch_properties" 2>&1 | tail -8) > $O/s5_pytest.log
(ROUNDS=2 timeout 600 bash tools/ab_bench.sh tools/ab/liblgx_old.so cylinder-pose-estimation_b200/liblgx.so tools/ab/liblgx_nst9.so tools/ab/liblgx_slp32.so tools/ab/liblgx_slp128.so 2>&1) > $O/s5_ab.log
cat $O/s5_pytest.log $O/s5_ab.log
