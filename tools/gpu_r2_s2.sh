#!/bin/bash
# Round-2 GPU session 2: cross-first accumulation order -- numerics (crossed paths), parity tests, bench
set -u
O=gpurun_out/${1:-r2s2}; mkdir -p $O
timeout 600 python tools/tc_numerics.py $O/numerics.json > $O/numerics.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided.json 2> $O/bench_guided.err
timeout 600 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided.json 2> $O/bench_unguided.err
grep "generic\|B16\|B48" $O/numerics.log; cat $O/pytest_gpu.log; cat $O/bench_guided.json; cat $O/bench_unguided.json
