#!/bin/bash
# quick A/B session: parity tests + guided bench (per-kernel) with an optional env switch.  usage: tools/gpu_quick.sh <outdir> [ENV=VAL for the B arm]
set -u
O=gpurun_out/${1:-quick}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q -x --tb=short > $O/pytest_gpu.log 2>&1
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided.json 2> $O/bench_guided.err
if [ -n "${2:-}" ]; then env "$2" timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided_B.json 2> $O/bench_guided_B.err; fi
O=$O python - <<'P'
import json,os
O=os.environ["O"]
for f in ("bench_guided.json","bench_guided_B.json"):
    if not os.path.exists(os.path.join(O,f)): continue
    d=json.load(open(os.path.join(O,f)))
    pk=d["roofline"]["per_kernel"]
    print(f, round(d["ms_per_step"],3), round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:v["ms_per_step"] for k,v in pk.items() if v["ms_per_step"]>0.3})
P
