#!/bin/bash
# steady-state DRAM traffic per launch (no cache flush between launches): shipped kernel vs the same without the state's L2 policy
mkdir -p gpurun_out
LEAN="--no-cpu-baseline --e2e-steps 0 --no-parity-check --no-policy-extras --no-packed"
for v in cur pol0; do
  for wl in hanabi_full_2p hanabi_small_2p; do
    HANABI_B200_LIB=$PWD/variants/$v.so timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:k_main -s 600 -c 16 --csv --log-file gpurun_out/r02ay_traffic_${v}_$wl.csv python bench.py --workload $wl --steps 20 --warmup 300 $LEAN > gpurun_out/r02ay_ncu_${v}_$wl.log 2>&1; echo "$v $wl rc=$?"
  done
done
python - <<'P'
import csv, glob
for f in sorted(glob.glob("gpurun_out/r02ay_traffic_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    agg = {}
    for r in rows:
        agg.setdefault(r[-3], []).append(float(r[-1].replace(",", "")))
    print(f, {k: round(sum(v) / len(v) / 1e6, 2) for k, v in agg.items()}, "launches", len(rows) // 3)
P
