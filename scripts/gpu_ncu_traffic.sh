#!/bin/bash
# DRAM traffic of the step kernel's launches in steady state WITHOUT ncu's cache flush between launches
# (--cache-control none): what the L2 policies save across steps is only visible this way
mkdir -p gpurun_out
LEAN="--no-cpu-baseline --e2e-steps 0 --no-parity-check --no-policy-extras --no-packed"
for wl in hanabi_full_2p hanabi_small_2p; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:k_main -s 600 -c 8 --csv --log-file gpurun_out/r02ax_traffic_$wl.csv python bench.py --workload $wl --steps 20 --warmup 300 $LEAN > gpurun_out/r02ax_ncu_$wl.log 2>&1; echo "$wl rc=$?"
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control all --clock-control none -k regex:k_main -s 600 -c 8 --csv --log-file gpurun_out/r02ax_traffic_flush_$wl.csv python bench.py --workload $wl --steps 20 --warmup 300 $LEAN > gpurun_out/r02ax_ncu_flush_$wl.log 2>&1; echo "$wl flush rc=$?"
done
python - <<'P'
import csv, glob
for f in sorted(glob.glob("gpurun_out/r02ax_traffic_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    agg = {}
    for r in rows:
        agg.setdefault(r[-3], []).append(float(r[-1].replace(",", "")))
    print(f, {k: round(sum(v) / len(v), 3) for k, v in agg.items()}, "launches", len(rows) // 3)
P
