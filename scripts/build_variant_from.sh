#!/bin/bash
# usage: build_variant_from.sh SRC_ROOT OUT.so [extra nvcc flags ...]  -- the whole library from another source tree
set -e
root=$1; out=$2; shift; shift
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=default"
cd $root/hanabi_b200/csrc
nvcc $FLAGS "$@" -shared -o $out hb_batch.cu hb_policy.cu hb_replay.cu hb_single.cc hb_host.cc
echo built $out
