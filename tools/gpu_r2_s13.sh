#!/bin/bash
# Round-2 GPU session 13: CUDA graph by default, side-stream overlap of the bond predictor's EdgeBlock tail
set -u
O=gpurun_out/${1:-r2s13}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > $O/pytest_gpu.log 2>&1
tail -6 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided_graph_ov1.json 2> $O/bench_guided_graph_ov1.err
MDB_OVERLAP=0 timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided_graph_ov0.json 2> $O/bench_guided_graph_ov0.err
timeout 600 python bench.py --no-graph --no-cpu-baseline > $O/bench_guided_eager_ov1.json 2> $O/bench_guided_eager_ov1.err
MDB_OVERLAP=0 timeout 600 python bench.py --no-graph --no-cpu-baseline > $O/bench_guided_eager_ov0.json 2> $O/bench_guided_eager_ov0.err
O=$O python - <<'P'
import json,sys,os
O=os.environ["O"]
for f in sorted(os.listdir(O)):
    if f.startswith("bench_") and f.endswith(".json"):
        try:
            d=json.load(open(os.path.join(O,f)))
            print(f, round(d["ms_per_step"],3), round(d["value"],3), "e2e", round(d["e2e"]["value"],3), d.get("gpu_launches"))
        except Exception as e: print(f, "ERR", e); print(open(os.path.join(O,f[:-5]+".err")).read()[-1500:])
P
