#!/bin/bash
# usage: variant_bench.sh lib1.so lib2.so ...   -> kernel ms + steps/s for each build of the library
# (environment such as HANABI_B200_L2_RESIDENT and BENCH_ARGS is passed through)
for lib in "$@"; do
  HANABI_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --e2e-steps 0 --no-policy-extras --no-packed --no-parity-check --warmup 100 --steps ${STEPS:-1000} $BENCH_ARGS 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'L2=${HANABI_B200_L2_RESIDENT:-0}', '$BENCH_ARGS', 'steps/s %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'])"
done
