#!/bin/bash
# A/B of library variants (variants/*.so given as arguments) against the shipped build, on three workloads
mkdir -p gpurun_out
for v in "$@"; do
  HANABI_B200_LIB=$PWD/$v timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lockstep and (full2 or small2 or full4)" 2>&1 | tail -1
done
STEPS=500 bash scripts/ab3.sh "hanabi_full_2p hanabi_small_2p hanabi_full_4p" "$@"
