#!/bin/bash
# Round-2 GPU session 12: split main / cross accumulators in the per-node backward chain -- numerics, parity, bench
set -u
O=gpurun_out/${1:-r2s12}; mkdir -p $O
timeout 600 python tools/tc_numerics.py $O/numerics.json > $O/numerics.log 2>&1
grep "B16\|B48" $O/numerics.log | grep "bwd=tc" | head -20
timeout 900 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu.log 2>&1
tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided.json 2> $O/bench_guided.err
O=$O python - <<'P'
import json,sys,os
O=os.environ["O"]
for f in sorted(os.listdir(O)):
    if f.startswith("bench_") and f.endswith(".json"):
        try:
            d=json.load(open(os.path.join(O,f)))
            pk=d["roofline"]["per_kernel"]
            print(f, round(d["ms_per_step"],3), {k:v["ms_per_step"] for k,v in pk.items() if v["ms_per_step"]>0.3})
        except Exception as e: print(f, "ERR", e); print(open(os.path.join(O,f[:-5]+".err")).read()[-1500:])
P
