#!/bin/bash
# last check of the round: smoke(), the full GPU suite, the driver's bench command and the reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_n1_driver.json 2> gpurun_out/final_bench_n1_driver.err; echo "bench rc=$?"
python - <<P
import json
for f in ('final_bench_reference', 'final_bench_n1_driver'):
  d=json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
  print(f, '%.4e' % d['value'], 'frac', (d.get('roofline') or {}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'))
P
