#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest tests/test_undistort.py -m gpu -x -q 2>&1 | tail -4) > $O/h_pytest.log
(timeout 100 python tools/undistort_prof.py 64 2>&1 | tail -2) > $O/h_und.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > $O/n2_bench.json 2> $O/n2_bench.err
cat $O/h_pytest.log $O/h_und.log; wc -l $O/n2_bench.json; cut -c1-400 $O/n2_bench.json; tail -2 $O/n2_bench.err
