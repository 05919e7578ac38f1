"""Top source lines by warp-stall samples (ncu source page): total samples per line with the
dominant stall reasons.  usage: ncu_stalls.py report.ncu-rep [top]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
tot = collections.Counter(); per = collections.defaultdict(collections.Counter); text = {}
cur = None; fn = None; kinds = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1]
    elif r and r[0] == "Function Name":
        if fn is None: fn = r[1]
        if r[1] != fn: break
    elif r and r[0] == "Line No":
        h = r; iA = h.index("Address"); iS = h.index("# Samples")
        kinds = [(c, h.index(c)) for c in h if c.startswith("stall_") and "Not Issued" not in c]
    elif cur and kinds and len(r) > 8 and r[iA] in ("", "-"):
        try:
            key = (cur.split('/')[-1], int(r[0])); text[key] = r[1].strip()[:70]
            tot[key] += int(r[iS] or 0)
            for k, i in kinds: per[key][k] += int(r[i] or 0)
        except ValueError: pass
allk = collections.Counter()
for key in per:
    for k, v in per[key].items(): allk[k] += v
T = sum(tot.values())
print("total samples", T, " by reason:", ", ".join("%s %.1f%%" % (k[6:], v / max(T,1) * 100) for k, v in allk.most_common(9)))
for key, v in tot.most_common(top):
    why = ", ".join("%s %d" % (k[6:], c) for k, c in per[key].most_common(3) if c)
    print(f"{v:6d} {v/max(T,1)*100:5.1f}%  {key[0]}:{key[1]:<4d} {text[key]}   [{why}]")
