"""Kernel time shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`): aggregates every launch from the
first sampler step on (first `node_init_kernel`), i.e. skips model construction / weight packing.
usage: python tools/launch_shares.py launches.csv out.csv "<command that was profiled>" """
import collections
import csv
import re
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
names, vals = [], []
for r in rows[1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n = re.sub(r"^void ", "", r[ki].split("(")[0].replace("<unnamed>::", ""))
    names.append(n)
    vals.append(v)
start = next((i for i, n in enumerate(names) if "node_init_kernel" in n), 0)
tot, cnt = collections.Counter(), collections.Counter()
for n, v in zip(names[start:], vals[start:]):
    key = n if not n.startswith(("native::", "at_cuda_detail::", "at::")) else "torch: " + n.split("<")[0]
    tot[key] += v
    cnt[key] += 1
T = sum(tot.values())
out = [f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised replays) of `{cmd}`:",
       f"# {len(names) - start} launches from the first sampler step on (launches {start}..{len(names) - 1}; model set-up skipped).",
       "# Compare SHARES with bench.py's roofline.per_kernel_ms, not absolute times.  unit: ns", "kernel,launches,total_ns,share"]
for k, v in tot.most_common():
    out.append(f"{k[:80]},{cnt[k]},{v:.0f},{v / T:.4f}")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:24]))
