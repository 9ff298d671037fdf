"""Split an ncu SASS source CSV into phases at the row threads' accumulator waits (the hot spin branches) and report
samples / executed instructions per phase.  usage: python tools/ncu_segments.py src_sass_X.csv.gz [min_wait_samples]"""
import csv, gzip, sys, io, collections
path = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 800
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(io.TextIOWrapper(op(path, "rb"))))
hdr = rows[1]; data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
S = ci["# Samples"]; SRC = ci["Source"]; IE = ci["Instructions Executed"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in data)
cuts = [i for i, r in enumerate(data) if "BRA" in r[SRC] and int(r[S] or 0) >= thr]
prev = 0
print(f"total samples {tot}")
for c in cuts + [len(data) - 1]:
    seg = data[prev:c]
    s = sum(int(r[S] or 0) for r in seg); ie = sum(int(r[IE] or 0) for r in seg)
    agg = collections.Counter()
    for r in seg:
        for i in stall_cols: agg[hdr[i][6:]] += int(r[i] or 0)
    ops = collections.Counter()
    for r in seg:
        t = r[SRC].split(); o = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
        ops[o] += int(r[IE] or 0)
    w = int(data[c][S] or 0)
    print(f"[{prev:6d},{c:6d}) n_sass {c-prev:5d} samples {s:6d} ({100*s/tot:4.1f}%) exec {ie/1e6:7.2f}M | then wait {w:5d} ({100*w/tot:4.1f}%) | "
          + ", ".join(f"{k} {100*v/max(s,1):.0f}%" for k, v in agg.most_common(4)) + " | " + ", ".join(f"{k} {v/1e6:.1f}M" for k, v in ops.most_common(6)))
    prev = c + 1
