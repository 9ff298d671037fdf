"""Split an ncu SASS source CSV at the kernel's clock64 phase stamps (CS2R ... SR_CLOCKLO) and report samples, executed
instructions, top stall reasons and top opcodes per phase.  usage: python tools/ncu_phases.py src_sass_X.csv.gz"""
import csv, gzip, sys, io, collections
path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(io.TextIOWrapper(op(path, "rb"))))
hdr = rows[1]; data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
S = ci["# Samples"]; SRC = ci["Source"]; IE = ci["Instructions Executed"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in data)
cuts = [i for i, r in enumerate(data) if "SR_CLOCK" in r[SRC]]
print(f"total samples {tot}, {len(cuts)} stamps, {len(data)} SASS instructions")
prev = 0
for k, c in enumerate(cuts + [len(data)]):
    seg = data[prev:c]
    s = sum(int(r[S] or 0) for r in seg); ie = sum(int(r[IE] or 0) for r in seg)
    agg = collections.Counter(); ops = collections.Counter(); opsamp = collections.Counter()
    for r in seg:
        for i in stall_cols: agg[hdr[i][6:]] += int(r[i] or 0)
        t = r[SRC].split(); o = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
        ops[o] += int(r[IE] or 0); opsamp[o] += int(r[S] or 0)
    print(f"phase {k:2d} sass[{prev:6d},{c:6d}) samples {s:6d} ({100*s/tot:4.1f}%) exec {ie/1e6:6.2f}M | "
          + ", ".join(f"{a} {100*v/max(s,1):.0f}%" for a, v in agg.most_common(4)) + " | "
          + ", ".join(f"{a} {100*v/max(s,1):.0f}%" for a, v in opsamp.most_common(6)))
    prev = c
