#!/bin/bash
# GPU session B: parity + bench after a kernel change (short)
set -u
O=gpurun_out/${1:-b}; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/pytest_gpu.log
timeout 400 python bench.py --workload guided --no-cpu-baseline > $O/bench_guided.json 2> $O/bench_guided.err
timeout 300 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided.json 2> $O/bench_unguided.err
cat $O/pytest_gpu.log; python - <<PY
import json
for f in ["$O/bench_guided.json","$O/bench_unguided.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],3), round(d["ms_per_step"],3), d["roofline"]["per_kernel_ms"])
    except Exception as e: print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
