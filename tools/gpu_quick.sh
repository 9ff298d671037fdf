#!/bin/bash
# quick A/B session: gradient parity subset + guided bench (per-kernel) + bwd16 phase table.  usage: tools/gpu_quick.sh <outdir>
set -u
O=gpurun_out/${1:-quick}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q -x --tb=short > $O/pytest_gpu.log 2>&1
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided.json 2> $O/bench_guided.err
timeout 300 python tools/tc_phase_times_bwd.py > $O/phase_bwd16.txt 2>&1
O=$O python - <<'P'
import json,os
O=os.environ["O"]
d=json.load(open(os.path.join(O,"bench_guided.json")))
pk=d["roofline"]["per_kernel"]
print(round(d["ms_per_step"],3), round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:v["ms_per_step"] for k,v in pk.items() if v["ms_per_step"]>0.3})
P
grep "epi\|total" $O/phase_bwd16.txt
