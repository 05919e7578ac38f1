#!/bin/bash
# the driver's scaling commands on one 8-GPU box: N = 1, 2, 4, 8 with --steps 20 --warmup 5 (final tree of round 2)
mkdir -p gpurun_out
nproc > gpurun_out/r02ar_host.txt; lscpu | grep -E "Model name|NUMA node" >> gpurun_out/r02ar_host.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02ar_scale_n1.json 2> gpurun_out/r02ar_scale_n1.err; echo "n1 rc=$?"
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29550+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02ar_scale_n$N.json 2> gpurun_out/r02ar_scale_n$N.err; echo "n$N rc=$?"
done
python - <<'P'
import json
for N in (1, 2, 4, 8):
    f = "r02ar_scale_n%d" % N
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        e=d["e2e"]
        print(f, "%.4e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "parity", d["parity_checked"], "e2e %.3e"%e["value"], "bytes link %.3e"%e["bytes_link"]["value"], "threads", e["host_policy_threads"], "rainbow", d.get("policies",{}).get("rainbow",{}).get("value"))
    except Exception as ex:
        print(f, "failed", ex); print(open("gpurun_out/%s.err"%f).read()[-800:])
P
cat gpurun_out/r02ar_host.txt
