#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "blur5 or seeded_size_sweep or frontend_small or frontend_cylinder or golden or config4 or strided" 2>&1 | tail -6) > $O/i_pytest.log
(ROUNDS=2 timeout 300 bash tools/ab_bench.sh tools/ab/liblgx_old.so cylinder-pose-estimation_b200/liblgx.so 2>&1) > $O/i_ab.log
cat $O/i_pytest.log $O/i_ab.log
