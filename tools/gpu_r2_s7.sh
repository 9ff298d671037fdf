#!/bin/bash
# Round-2 GPU session 7: two-CTAs-per-SM BondFFN forward kernel -- parity, A/B bench, phase times
set -u
O=gpurun_out/${1:-r2s7}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > $O/pytest_gpu.log 2>&1
tail -15 $O/pytest_gpu.log
for v in 1 0; do
  MDB_TC_FFN2=$v timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided_ffn2_$v.json 2> $O/bench_guided_ffn2_$v.err
  MDB_TC_FFN2=$v timeout 600 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided_ffn2_$v.json 2> $O/bench_unguided_ffn2_$v.err
done
timeout 300 python tools/tc_phase_times_ffn.py > $O/phase_ffn.txt 2>&1
python - <<'P'
import json,sys,os
O=os.environ.get("O","gpurun_out/r2s7")
for f in sorted(os.listdir(O)):
    if f.startswith("bench_") and f.endswith(".json"):
        try:
            d=json.load(open(os.path.join(O,f)))
            pk=d["roofline"]["per_kernel"]
            print(f, round(d["ms_per_step"],3), {k:v["ms_per_step"] for k,v in pk.items() if k.startswith("tc_")})
        except Exception as e: print(f, "ERR", e)
P
cat $O/phase_ffn.txt
