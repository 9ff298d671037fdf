#!/bin/bash
# Round-2 GPU session 16: denoiser node kernel split (mid in the chain, pre beside the edge kernel)
set -u
O=gpurun_out/${1:-r2s16}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > $O/pytest_gpu.log 2>&1
tail -6 $O/pytest_gpu.log
for ov in 1 0; do
MDB_OVERLAP=$ov timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided_ov$ov.json 2> $O/bench_guided_ov$ov.err
MDB_OVERLAP=$ov timeout 600 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided_ov$ov.json 2> $O/bench_unguided_ov$ov.err
MDB_OVERLAP=$ov timeout 600 python bench.py --simple --no-cpu-baseline > $O/bench_simple_ov$ov.json 2> $O/bench_simple_ov$ov.err
MDB_OVERLAP=$ov timeout 600 python bench.py --workload train_fwd --no-cpu-baseline > $O/bench_trainfwd_ov$ov.json 2> $O/bench_trainfwd_ov$ov.err
done
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
