#!/bin/bash
# the driver's scaling commands on one 8-GPU box: N = 1 and N = 8, --steps 20 --warmup 5
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_scale_n1_driver.json 2> gpurun_out/r02_scale_n1_driver.err; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_scale_n8_driver.json 2> gpurun_out/r02_scale_n8_driver.err; echo "n8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 1000 --warmup 50 --e2e-steps 0 --no-packed --no-policy-extras > gpurun_out/r02_scale_n8_long.json 2> gpurun_out/r02_scale_n8_long.err; echo "n8 long rc=$?"
python - <<'P'
import json
for f in ("r02_scale_n1_driver","r02_scale_n8_driver","r02_scale_n8_long"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "%.4e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "parity", d["parity_checked"], "e2e", d["e2e"] and "%.3e"%d["e2e"]["value"], [round(x,4) for x in d["ranks"]["ms_per_step"]], d["ranks"]["stats_allreduce_us"], d.get("policies",{}).get("rainbow",{}).get("value"))
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
P
