"""Summarise an ncu report: key metrics + executed warp-instructions per source line."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'dram__bytes.sum.per_second','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__sass_branch_targets_threads_divergent.sum',
 'dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','sm__inst_executed_pipe_lsu.sum','smsp__inst_executed_pipe_alu.sum','smsp__inst_executed_pipe_fma.sum']
r = rows[2]
print('kernel', r[hdr.index('Kernel Name')][:70])
for w in want:
    if w in hdr: print(f'  {w:66s} {units[hdr.index(w)]:10s} {r[hdr.index(w)]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = collections.Counter(); thr = collections.Counter(); text = {}
cur = None; seen_kernel = 0; fn = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1]
    elif r and r[0] == "Function Name":
        if fn is None: fn = r[1]
        if r[1] != fn: break
    elif r and r[0] == "Line No": h = r; iA = h.index("Address"); iI = h.index("Instructions Executed"); iT = h.index("Thread Instructions Executed")
    elif cur and len(r) > 8 and r[iA] in ("", "-"):
        try:
            k = (cur.split('/')[-1], int(r[0])); agg[k] += int(r[iI]); thr[k] += int(r[iT]); text[k] = r[1].strip()[:90]
        except ValueError: pass
tot = sum(agg.values()); print('total warp-instr', tot)
for k, v in agg.most_common(top):
    print(f'{v:>10d} {v/tot*100:5.1f}% lanes {thr[k]/max(v,1):4.1f}  {k[0]}:{k[1]:<4d} {text[k]}')
