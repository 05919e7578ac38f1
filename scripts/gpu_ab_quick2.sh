#!/bin/bash
# evict-last on the packed state: where does it pay?  5p / 4p shards, unsharded 2p, 2^22 and 2^18 games
for lib in variants/base.so variants/st.so variants/base.so variants/st.so; do
  BENCH_ARGS="--workload hanabi_full_5p" STEPS=400 bash scripts/variant_bench.sh $lib
  BENCH_ARGS="--workload hanabi_small_4p" STEPS=400 bash scripts/variant_bench.sh $lib
  BENCH_ARGS="--shards 1" STEPS=400 bash scripts/variant_bench.sh $lib
  BENCH_ARGS="--games 4194304" STEPS=150 bash scripts/variant_bench.sh $lib
  BENCH_ARGS="--games 262144" STEPS=1000 bash scripts/variant_bench.sh $lib
done
