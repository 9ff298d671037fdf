#!/bin/bash
# Round-2 GPU session 3: cross-first order only in the bond predictor's forward kernels -- numerics, tests, bench
set -u
O=gpurun_out/${1:-r2s3}; mkdir -p $O
timeout 600 python tools/tc_numerics.py $O/numerics.json > $O/numerics.log 2>&1
MDB_LIB_VARIANT=exactsig timeout 600 python tools/tc_numerics.py $O/numerics_exactsig.json > $O/numerics_exactsig.log 2>&1
MDB_TC_NODE=0 timeout 600 python tools/tc_numerics.py $O/numerics_ffnode.json > $O/numerics_ffnode.log 2>&1
MDB_CROSS_FIRST=1 timeout 600 python tools/tc_numerics.py $O/numerics_xfall.json > $O/numerics_xfall.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25) > $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided.json 2> $O/bench_guided.err
timeout 600 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided.json 2> $O/bench_unguided.err
for f in numerics numerics_exactsig numerics_ffnode numerics_xfall; do echo "== $f"; grep "generic\|B16\|B48" $O/$f.log | grep -v "K64\|K128"; done
cat $O/pytest_gpu.log; cat $O/bench_guided.json; cat $O/bench_unguided.json
