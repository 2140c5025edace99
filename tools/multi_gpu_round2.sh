#!/bin/bash
# [AFF=1] multi_gpu_round2.sh N: the round-2 multi-GPU record on one box (run under gpurun --gpus N): topology, copy-only probe (with and
# without CPU affinity), the default bench (config 3, weak scaling) and the two sharded jobs (configs 4 and 5, strong scaling).
N=$1
O=gpurun_out/mg_n$N
mkdir -p $O
run() { if [ "$N" = 1 ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 "$@"; fi; }
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | grep -i -E "numa|model name|^cpu\(s\)" >> $O/topo.txt
run tools/copy_probe.py > $O/copy_probe.json 2> $O/copy_probe.err
[ -n "$AFF" ] && LGX_AFFINITY=1 run tools/copy_probe.py > $O/copy_probe_affinity.json 2>> $O/copy_probe.err
run bench.py --gpus $N --steps 10 --warmup 3 --check 0 --no-cpu > $O/bench_cfg3.json 2> $O/bench_cfg3.err
[ -n "$AFF" ] && run bench.py --gpus $N --steps 10 --warmup 3 --check 0 --no-cpu --affinity > $O/bench_cfg3_affinity.json 2>> $O/bench_cfg3.err
run bench.py --gpus $N --config 4 --steps 3 --warmup 1 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
run bench.py --gpus $N --config 5 --steps 3 --warmup 1 > $O/bench_cfg5.json 2> $O/bench_cfg5.err
for f in copy_probe copy_probe_affinity; do cat $O/$f.json; done
for f in bench_cfg3 bench_cfg3_affinity bench_cfg4 bench_cfg5; do python -c "
import json,sys
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print('$f', 'N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],2))
except Exception as e: print('$f FAILED', e)
"; done
for f in $O/*.err; do tail -n 2 $f; done | tail -n 12
