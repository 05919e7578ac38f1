#!/bin/bash
# the driver's N=2 command on one 2-GPU box + the gloo/NCCL-free checks
mkdir -p gpurun_out
nproc > gpurun_out/r02y_nproc.txt; numactl -H 2>/dev/null | head -5 >> gpurun_out/r02y_nproc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02y_bench_n2_driver.json 2> gpurun_out/r02y_bench_n2_driver.err; echo "n2 rc=$?"
python - <<'P'
import json
for f in ("r02y_bench_n2_driver",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        e=d["e2e"]
        print(f, "%.4e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "parity", d["parity_checked"], "e2e %.3e"%e["value"], "bytes link %.3e"%e["bytes_link"]["value"], e["host_policy_threads"], d.get("policies",{}).get("rainbow",{}).get("value"))
    except Exception as ex:
        print(f, "failed", ex); print(open("gpurun_out/%s.err"%f).read()[-800:])
P
cat gpurun_out/r02y_nproc.txt
