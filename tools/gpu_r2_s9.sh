#!/bin/bash
# Round-2 GPU session 9: mbarrier try_wait suspend hint (A/B: MDB_LIB_VARIANT=nohint) x two-CTA BondFFN forward (MDB_TC_FFN2)
set -u
O=gpurun_out/${1:-r2s9}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --tb=short > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
for lib in "" nohint; do for v in 1 0; do
  MDB_LIB_VARIANT=$lib MDB_TC_FFN2=$v timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided_${lib:-hint}_ffn2_$v.json 2> $O/bench_guided_${lib:-hint}_ffn2_$v.err
done; done
timeout 300 python tools/tc_phase_times_ffn.py > $O/phase_ffn.txt 2>&1
O=$O python - <<'P'
import json,sys,os
O=os.environ["O"]
for f in sorted(os.listdir(O)):
    if f.startswith("bench_") and f.endswith(".json"):
        try:
            d=json.load(open(os.path.join(O,f)))
            pk=d["roofline"]["per_kernel"]
            print(f, round(d["ms_per_step"],3), {k:v["ms_per_step"] for k,v in pk.items() if v["ms_per_step"]>0.3})
        except Exception as e: print(f, "ERR", e)
P
cat $O/phase_ffn.txt
