#!/bin/bash
# usage: build_variant.sh OUT.so [extra nvcc flags for hb_batch.cu ...]
# Rebuilds only hb_batch.cu (the step kernel) and links it with cached objects of the other sources.
set -e
cd "$(dirname "$0")/.."
out=$1; shift
cache=/tmp/hb_variant_objs
mkdir -p $cache
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=default"
# POLICY_FLAGS: extra flags for hb_policy.cu (forces its rebuild into a private object)
for src in hb_policy.cu hb_replay.cu hb_single.cc hb_host.cc; do
  obj=$cache/${src%.*}.o
  if [ ! -f $obj ] || [ hanabi_b200/csrc/$src -nt $obj ]; then nvcc $FLAGS -c hanabi_b200/csrc/$src -o $obj; fi
done
nvcc $FLAGS "$@" -c hanabi_b200/csrc/hb_batch.cu -o $cache/hb_batch_$$.o
pol=$cache/hb_policy.o
if [ -n "$POLICY_FLAGS" ]; then pol=$cache/hb_policy_$$.o; nvcc $FLAGS $POLICY_FLAGS -c hanabi_b200/csrc/hb_policy.cu -o $pol; fi
nvcc -shared -o $out $cache/hb_batch_$$.o $pol $cache/hb_replay.o $cache/hb_single.o $cache/hb_host.o
rm -f $cache/hb_batch_$$.o
echo built $out
