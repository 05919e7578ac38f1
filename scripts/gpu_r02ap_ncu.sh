#!/bin/bash
# ncu --set full of the final step kernel: one launch over 2^20 games, one of the default bench's two 2^19-game launches,
# the Small 2p instantiation, and the launch list of the default bench
mkdir -p gpurun_out
LEAN="--no-cpu-baseline --e2e-steps 0 --no-parity-check --no-policy-extras --no-packed"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_main -s 305 -c 1 -f -o gpurun_out/r02ap_step_full2p python bench.py --steps 20 --warmup 300 --shards 1 $LEAN > gpurun_out/r02ap_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_main -s 610 -c 1 -f -o gpurun_out/r02ap_step_full2p_half python bench.py --steps 20 --warmup 300 $LEAN > gpurun_out/r02ap_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_main -s 305 -c 1 -f -o gpurun_out/r02ap_step_small2p python bench.py --workload hanabi_small_2p --steps 20 --warmup 300 --shards 1 $LEAN > gpurun_out/r02ap_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02ap_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-policy-extras --no-packed > gpurun_out/r02ap_ncu4.log 2>&1; echo "ncu4 rc=$?"
ls -la gpurun_out/r02ap_*
