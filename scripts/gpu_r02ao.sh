#!/bin/bash
# final kernel of round 2 (per-view L2 policies): full GPU suite, the five workloads, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02ao_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ao_pytest.log
for w in hanabi_full_4p hanabi_full_5p hanabi_small_2p hanabi_small_4p; do
  timeout 300 python bench.py --workload $w --steps 1000 --warmup 100 --no-cpu-baseline --e2e-steps 0 --no-policy-extras --no-packed > gpurun_out/r02ao_bench_$w.json 2> gpurun_out/r02ao_bench_$w.err; echo "$w rc=$?"
done
timeout 900 python bench.py > gpurun_out/r02ao_bench_n1.json 2> gpurun_out/r02ao_bench_n1.err; echo "bench rc=$?"
python - <<P
import json
for f in ('bench_hanabi_full_4p','bench_hanabi_full_5p','bench_hanabi_small_2p','bench_hanabi_small_4p','bench_n1'):
  try:
    d=json.loads(open('gpurun_out/r02ao_%s.json' % f).read().strip().splitlines()[-1])
    e=d.get('e2e') or {}
    print(f, '%.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'parity', d['parity_checked'], 'e2e', e.get('value'))
  except Exception as ex:
    print(f, 'failed', ex)
P
