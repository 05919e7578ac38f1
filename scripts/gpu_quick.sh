#!/bin/bash
# quick GPU check: tests (stop at first failure) + lean bench lines for the given workloads
# usage: gpu_quick.sh TAG [workload ...]
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
for w in "$@"; do
  timeout 300 python bench.py --workload $w --steps 1000 --warmup 100 --no-cpu-baseline --e2e-steps 0 --no-policy-extras --no-packed > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "$w rc=$?"
  python - <<P
import json
try:
  d=json.loads(open('gpurun_out/${TAG}_bench_$w.json').read().strip().splitlines()[-1])
  print('$w', '%.4e'%d['value'], 'ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'parity', d['parity_checked'], 'eplen', round(d['episodes']['mean_length'],3))
except Exception as e:
  print('$w failed', e); print(open('gpurun_out/${TAG}_bench_$w.err').read()[-1500:])
P
done
