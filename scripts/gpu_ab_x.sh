#!/bin/bash
# parity subset with each variant, then A/B on five workloads: gpu_ab_x.sh base.so variant.so ...
for v in "${@:2}"; do
  HANABI_B200_LIB=$PWD/$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_env.py -m gpu -x -q 2>&1 | tail -1
done
STEPS=500 bash scripts/ab3.sh "hanabi_full_2p hanabi_small_2p hanabi_full_4p hanabi_full_5p" "$@"
