#!/bin/bash
# round 2, first GPU call: GPU tests, the driver's own bench command, the reference arm, the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_bench_n1_driver.json 2> gpurun_out/r02a_bench_n1_driver.err; echo "bench20 rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02a_bench_n1_driver.err
head -c 1500 gpurun_out/r02a_bench_n1_driver.json
