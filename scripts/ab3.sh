#!/bin/bash
# usage: ab3.sh "workload ..." lib...  -> bench every build on every workload, twice, interleaved
WLS=$1; shift
for wl in $WLS; do for rep in 1 2; do for lib in "$@"; do BENCH_ARGS="--workload $wl" STEPS=${STEPS:-600} bash scripts/variant_bench.sh $lib; done; done; done
