#!/bin/bash
# A/B: alternate libraries (default: tools/ab/liblgx_old.so and the current one), ROUNDS rounds, same box
LIBS=${@:-tools/ab/liblgx_old.so cylinder-pose-estimation_b200/liblgx.so}
for r in $(seq ${ROUNDS:-3}); do
  for lib in $LIBS; do
    LGX_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu --check 0 2>/dev/null | LIBNAME=$lib python -c "
import sys,json,os; d=json.loads(sys.stdin.read()); r=d['roofline']; s=r['kernel_ms_share']; f=d['ms_per_step']/256*1e3
print(os.environ['LIBNAME'][-14:], 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'us/frame:', {k: round(v*f,1) for k,v in s.items()})"
  done
done
