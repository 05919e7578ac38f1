"""Top source lines by warp-stall samples of given kinds (ncu source page)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; kinds = sys.argv[2].split(','); top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = {k: collections.Counter() for k in kinds}; text = {}
cur = None; fn = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1]
    elif r and r[0] == "Function Name":
        if fn is None: fn = r[1]
        if r[1] != fn: break
    elif r and r[0] == "Line No": h = r; iA = h.index("Address"); idx = {k: h.index(k) for k in kinds}
    elif cur and len(r) > 8 and r[iA] in ("", "-"):
        try:
            key = (cur.split('/')[-1], int(r[0])); text[key] = r[1].strip()[:80]
            for k in kinds: agg[k][key] += int(r[idx[k]] or 0)
        except ValueError: pass
for k in kinds:
    tot = sum(agg[k].values()); print(f"== {k}: {tot} samples")
    for key, v in agg[k].most_common(top):
        print(f"  {v:6d} {v/max(tot,1)*100:5.1f}%  {key[0]}:{key[1]:<4d} {text[key]}")
