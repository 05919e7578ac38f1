#!/bin/bash
# usage: ab2.sh lib...  -> bench each build on the headline workload (twice) and on Small 2p / Full 4p
for rep in 1 2; do for lib in "$@"; do bash scripts/variant_bench.sh $lib; done; done
for wl in hanabi_full_4p hanabi_small_2p; do
  for lib in "$@"; do BENCH_ARGS="--workload $wl" STEPS=500 bash scripts/variant_bench.sh $lib; done
done
