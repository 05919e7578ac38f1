#!/bin/bash
# wide hands (6-8 cards) + vectorised host policy: full GPU suite, then the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02v_pytest.log
timeout 900 python bench.py --no-policy-extras --python-seconds 0 > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02v_bench_n1.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02v_bench_n1.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.4e frac %.4f' % (d['value'], d['roofline']['frac']))
print('e2e packed link %.4e  ms %.3f  pcie %s' % (e['value'], e['ms_per_step'], e['pcie']))
print('e2e bytes link  %.4e  ms %.3f' % (e['bytes_link']['value'], e['bytes_link']['ms_per_step']))
print('single', e['single_call'])
P
