"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): the tcgen05 / TMEM / bulk-copy evidence the
profiling guide asks for -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), FFMA, RED, MUFU.  usage: python tools/sass_hist.py [lib.so] > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "moldiff_b200",
                                                         "libmoldiff_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA", "FADD", "FMUL", "F2FP", "MUFU", "RED", "ATOMG", "LDG", "STG", "LDS", "STS",
         "BAR", "USETMAXREG"]
kern, hist = None, {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern).split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
        hist[kern]["_total"] += 1
print(f"# cuobjdump -sass {os.path.basename(lib)} ({os.path.getsize(lib)} bytes): static SASS instruction counts per kernel")
print("# " + " ".join(f"{w:>8s}" for w in ["total"] + WATCH) + "  kernel")
for k, h in sorted(hist.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 100000 - kv[1]["_total"]):
    if h["_total"] < 50:
        continue
    print("  " + " ".join(f"{h[w]:8d}" for w in ["_total"] + WATCH) + "  " + k[:90])
