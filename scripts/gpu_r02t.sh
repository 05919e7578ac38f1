#!/bin/bash
# packed host link: the two host-path GPU tests, then the default bench line (both link forms in e2e)
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-400 > gpurun_out/r02t_cpu.txt
timeout 600 python -m pytest tests/test_gpu_env.py -m gpu -x -q -k "host" > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02t_pytest.log
timeout 900 python bench.py --no-policy-extras --python-seconds 0 > gpurun_out/r02t_bench_n1.json 2> gpurun_out/r02t_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02t_bench_n1.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02t_bench_n1.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.4e frac %.4f' % (d['value'], d['roofline']['frac']))
print('e2e packed link %.4e  ms %.3f  pcie %s' % (e['value'], e['ms_per_step'], e['pcie']))
print('e2e bytes link  %.4e  ms %.3f' % (e['bytes_link']['value'], e['bytes_link']['ms_per_step']))
print('single', e['single_call'])
print('packed_obs e2e', d.get('packed_obs',{}).get('e2e'))
P
