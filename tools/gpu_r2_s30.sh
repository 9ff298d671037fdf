#!/bin/bash
# Round-2 GPU session 30: packed weight slots (several K stages of a narrow GEMM per ring slot) -- selftest, parity, bench, phase tables
set -u
O=gpurun_out/${1:-r2s30}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > $O/pytest_gpu.log 2>&1
tail -6 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided.json 2> $O/bench_guided.err
timeout 600 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided.json 2> $O/bench_unguided.err
timeout 300 python tools/tc_phase_times_bwd.py > $O/phase_bwd16.txt 2>&1
timeout 300 python tools/tc_phase_times_ffn.py > $O/phase_ffn.txt 2>&1
O=$O python - <<'P'
import json,sys,os
O=os.environ["O"]
for f in sorted(os.listdir(O)):
    if f.startswith("bench_") and f.endswith(".json"):
        try:
            d=json.load(open(os.path.join(O,f)))
            pk=d["roofline"]["per_kernel"]
            print(f, round(d["ms_per_step"],3), round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:v["ms_per_step"] for k,v in pk.items() if v["ms_per_step"]>0.3})
        except Exception as e: print(f, "ERR", e); print(open(os.path.join(O,f[:-5]+".err")).read()[-1500:])
P
grep "wait" $O/phase_bwd16.txt; grep "wait\|total" $O/phase_ffn.txt
