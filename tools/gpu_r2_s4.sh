#!/bin/bash
# Round-2 GPU session 4: FFMA node path for the bond predictor, with / without the cross-first order; full test log; new bench modes
set -u
O=gpurun_out/${1:-r2s4}; mkdir -p $O
timeout 600 python tools/tc_numerics.py $O/numerics.json > $O/numerics.log 2>&1
MDB_CROSS_FIRST=0 timeout 600 python tools/tc_numerics.py $O/numerics_noxf.json > $O/numerics_noxf.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided.json 2> $O/bench_guided.err
MDB_CROSS_FIRST=0 timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided_noxf.json 2> $O/bench_guided_noxf.err
timeout 600 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided.json 2> $O/bench_unguided.err
timeout 600 python bench.py --simple --no-cpu-baseline > $O/bench_simple.json 2> $O/bench_simple.err
timeout 600 python bench.py --simple --graph --no-cpu-baseline > $O/bench_simple_graph.json 2> $O/bench_simple_graph.err
timeout 600 python bench.py --graph --no-cpu-baseline > $O/bench_guided_graph.json 2> $O/bench_guided_graph.err
timeout 600 python bench.py --workload train_fwd --no-cpu-baseline > $O/bench_train_fwd.json 2> $O/bench_train_fwd.err
for f in numerics numerics_noxf; do echo "== $f"; grep "B16\|B48" $O/$f.log | grep "fwd=tc"; done
tail -40 $O/pytest_gpu.log
for f in guided guided_noxf unguided simple simple_graph guided_graph train_fwd; do echo "== $f"; head -c 600 $O/bench_$f.json; echo; tail -3 $O/bench_$f.err; done
