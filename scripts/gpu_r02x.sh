#!/bin/bash
# final-tree evidence, one GPU: the driver's own commands (N=1 bench, reference arm), then the default bench line
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02x_bench_n1_driver.json 2> gpurun_out/r02x_bench_n1_driver.err; echo "driver bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02x_bench_reference.json 2> gpurun_out/r02x_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r02x_bench_n1.json 2> gpurun_out/r02x_bench_n1.err; echo "bench rc=$?"
python - <<P
import json
for f in ('r02x_bench_n1_driver', 'r02x_bench_reference', 'r02x_bench_n1'):
  try:
    d=json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
    e=d.get('e2e') or {}
    print(f, '%.4e' % d['value'], 'ms', d.get('ms_per_step'), 'frac', (d.get('roofline') or {}).get('frac'), 'e2e', e.get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'rainbow', ((d.get('policies') or {}).get('rainbow') or {}).get('value'))
  except Exception as ex:
    print(f, 'failed', ex); print(open('gpurun_out/%s.err' % f).read()[-600:])
P
