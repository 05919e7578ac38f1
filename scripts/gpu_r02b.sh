#!/bin/bash
# round 2: 2-GPU check of the bench's multi-rank window + cross-GPU shard parity test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k "two_gpu or sharded" > gpurun_out/r02b_pytest2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02b_pytest2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b_bench_n2_driver.json 2> gpurun_out/r02b_bench_n2_driver.err; echo "bench n2 rc=$?"
tail -c 400 gpurun_out/r02b_bench_n2_driver.err
head -c 600 gpurun_out/r02b_bench_n2_driver.json
