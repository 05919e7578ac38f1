#!/bin/bash
# round 2: new GPU tests (fixtures), Hanabi-Small bench lines and the first ncu capture of the Small-2p step kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02c_pytest.log
LEAN="--no-cpu-baseline --e2e-steps 0 --no-parity-check --no-policy-extras --no-packed"
for w in hanabi_small_2p hanabi_small_4p hanabi_full_4p hanabi_full_5p; do
  timeout 300 python bench.py --workload $w --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 0 --no-policy-extras --no-packed > gpurun_out/r02c_bench_$w.json 2> gpurun_out/r02c_bench_$w.err; echo "$w rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_main -s 305 -c 1 -f -o gpurun_out/r02c_small2p python bench.py --workload hanabi_small_2p --steps 20 --warmup 300 --shards 1 $LEAN > gpurun_out/r02c_ncu_small2p.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02c_small2p.ncu-rep --page raw --csv > gpurun_out/r02c_small2p_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
