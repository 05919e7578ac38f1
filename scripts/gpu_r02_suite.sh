#!/bin/bash
# round 2 evidence run on one GPU: bench lines of every workload and policy mode, the game-count
# sweep of SURVEY 8d, launch list + ncu --set full captures of the step kernels
T=${1:-r02}
mkdir -p gpurun_out
LEAN="--no-cpu-baseline --e2e-steps 0 --no-policy-extras --no-packed"
for w in hanabi_full_4p hanabi_full_5p hanabi_small_2p hanabi_small_4p; do
  timeout 300 python bench.py --workload $w --steps 1000 --warmup 100 $LEAN > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "$w rc=$?"
done
for g in 65536 262144 4194304; do
  timeout 300 python bench.py --games $g --steps 1000 --warmup 100 $LEAN > gpurun_out/${T}_sweep_$g.json 2> gpurun_out/${T}_sweep_$g.err; echo "sweep $g rc=$?"
done
timeout 300 python bench.py --policy rainbow --games 262144 --mlp-dtype bf16 --steps 300 --warmup 30 $LEAN > gpurun_out/${T}_bench_rainbow_bf16.json 2> gpurun_out/${T}_bench_rainbow_bf16.err; echo "rainbow rc=$?"
timeout 300 python bench.py --policy rainbow --record --games 262144 --mlp-dtype bf16 --steps 300 --warmup 30 $LEAN > gpurun_out/${T}_bench_actor_loop_bf16.json 2> gpurun_out/${T}_bench_actor_loop_bf16.err; echo "actor loop rc=$?"
timeout 300 python bench.py --policy ppo --workload hanabi_small_4p --games 1048576 --mlp-dtype bf16 --steps 300 --warmup 30 $LEAN > gpurun_out/${T}_bench_ppo_small_4p_bf16.json 2> gpurun_out/${T}_bench_ppo.err; echo "ppo rc=$?"
timeout 300 python bench.py --policy adhoc --workload hanabi_full_4p --games 65536 --mlp-dtype bf16 --steps 300 --warmup 30 $LEAN > gpurun_out/${T}_bench_adhoc_65536_bf16.json 2> gpurun_out/${T}_bench_adhoc.err; echo "adhoc 65536 rc=$?"
timeout 300 python bench.py --policy adhoc --workload hanabi_full_4p --games 8192 --mlp-dtype bf16 --steps 800 --warmup 40 $LEAN > gpurun_out/${T}_bench_adhoc_8192_bf16.json 2>> gpurun_out/${T}_bench_adhoc.err; echo "adhoc 8192 (8 steps per graph) rc=$?"
timeout 300 python bench.py --policy adhoc --workload hanabi_full_4p --games 8192 --mlp-dtype bf16 --steps 800 --warmup 40 --graph-steps 1 $LEAN > gpurun_out/${T}_bench_adhoc_8192_bf16_graph1.json 2>> gpurun_out/${T}_bench_adhoc.err; echo "adhoc 8192 (1 step per graph) rc=$?"
# launch list of the default bench (nothing but the step kernel in the timed region)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 6 --warmup 3 $LEAN --no-parity-check > gpurun_out/${T}_ncu_launches.log 2>&1; echo "launch list rc=$?"
NCU="--no-cpu-baseline --e2e-steps 0 --no-parity-check --no-policy-extras --no-packed"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_main -s 305 -c 1 -f -o gpurun_out/${T}_step_full2p python bench.py --steps 20 --warmup 300 --shards 1 $NCU > gpurun_out/${T}_ncu_full2p.log 2>&1; echo "ncu full2p rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_main -s 610 -c 1 -f -o gpurun_out/${T}_step_full2p_half python bench.py --steps 20 --warmup 300 $NCU > gpurun_out/${T}_ncu_full2p_half.log 2>&1; echo "ncu full2p half rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_main -s 305 -c 1 -f -o gpurun_out/${T}_step_full4p python bench.py --workload hanabi_full_4p --steps 20 --warmup 300 --shards 1 $NCU > gpurun_out/${T}_ncu_full4p.log 2>&1; echo "ncu full4p rc=$?"
for r in full2p full2p_half full4p; do ncu -i gpurun_out/${T}_step_$r.ncu-rep --page raw --csv > gpurun_out/${T}_step_${r}_raw.csv 2>/dev/null; done
ls gpurun_out | grep ${T}_ | wc -l
