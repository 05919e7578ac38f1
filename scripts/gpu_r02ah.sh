#!/bin/bash
# evict-last state per view: full GPU suite, A/B against the previous build on all workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02ah_pytest.log
STEPS=500 bash scripts/ab3.sh "hanabi_full_2p hanabi_small_2p hanabi_small_4p hanabi_full_4p hanabi_full_5p" variants/base.so hanabi_b200/libpyhanabi.so
