#!/bin/bash
# usage: gpu_ncu.sh TAG WORKLOAD : ncu --set full of one steady-state launch of the step kernel (one launch over all games)
TAG=$1; W=$2
mkdir -p gpurun_out
LEAN="--no-cpu-baseline --e2e-steps 0 --no-parity-check --no-policy-extras --no-packed"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_main -s 305 -c 1 -f -o gpurun_out/${TAG}_${W} python bench.py --workload $W --steps 20 --warmup 300 --shards 1 $LEAN > gpurun_out/${TAG}_ncu_${W}.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_${W}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${W}_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_${W}*
