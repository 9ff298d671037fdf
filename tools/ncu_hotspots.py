"""Top stall hotspots of an ncu source-page CSV (SASS view): usage  python tools/ncu_hotspots.py src_sass_X.csv.gz [N]"""
import csv, gzip, sys, io, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(io.TextIOWrapper(op(path, "rb"))))
hdr = rows[1]; data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
S = ci["# Samples"]; SRC = ci["Source"]; IE = ci["Instructions Executed"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in data)
print(f"instructions {len(data)}  samples {tot}  inst_executed {sum(int(r[IE] or 0) for r in data)}")
agg = collections.Counter()
for r in data:
    for i in stall_cols: agg[hdr[i]] += int(r[i] or 0)
print("stall totals:", ", ".join(f"{k[6:]} {v} ({100*v/max(tot,1):.0f}%)" for k, v in agg.most_common(8)))
# by opcode
byop = collections.Counter(); cnt = collections.Counter()
for r in data:
    t = r[SRC].split()
    opc = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    opc = opc.split(".")[0]
    byop[opc] += int(r[S] or 0); cnt[opc] += int(r[IE] or 0)
print("samples by opcode:", ", ".join(f"{k} {v} ({100*v/max(tot,1):.0f}%; {cnt[k]} exec)" for k, v in byop.most_common(14)))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][S] or 0))[:topn]
for i in sorted(idx):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:6d} {int(r[S]):7d} {100*int(r[S])/tot:5.1f}%  {r[SRC].strip()[:90]:90s} {st}")
