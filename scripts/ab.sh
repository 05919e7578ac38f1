#!/bin/bash
# usage: ab.sh variant.so  -> parity tests with the variant, then A/B bench against the shipped build
v=$1
HANABI_B200_LIB=$PWD/$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_env.py -x -q 2>&1 | tail -4
for lib in ${BASE:-hanabi_b200/libpyhanabi.so} $v ${BASE:-hanabi_b200/libpyhanabi.so} $v; do bash scripts/variant_bench.sh $lib; done
for wl in hanabi_full_4p hanabi_full_5p hanabi_small_2p hanabi_small_4p; do
  for lib in ${BASE:-hanabi_b200/libpyhanabi.so} $v; do BENCH_ARGS="--workload $wl" STEPS=500 bash scripts/variant_bench.sh $lib; done
done
