#!/bin/bash
# Round-2 GPU session 1: tensor-core numerics study, crossed gradient paths, parity tests, bench after the per-CTA scratch change
set -u
O=gpurun_out/${1:-r2s1}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 600 python tools/tc_numerics.py $O/numerics.json > $O/numerics.log 2>&1
MDB_LIB_VARIANT=exactsig timeout 600 python tools/tc_numerics.py $O/numerics_exactsig.json > $O/numerics_exactsig.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_guided.json 2> $O/bench_guided.err
k=tc_nodeblock_bwd16
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 -c 1 -f -o $O/full_$k \
   python bench.py --workload guided --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_$k.log 2>&1
ncu -i $O/full_$k.ncu-rep --page raw --csv > $O/raw_$k.csv 2>/dev/null
ncu -i $O/full_$k.ncu-rep --page details > $O/details_$k.txt 2>/dev/null
rm -f $O/full_$k.ncu-rep
tail -30 $O/numerics.log; grep crossed -A0 $O/numerics_exactsig.log | head -0; tail -17 $O/numerics_exactsig.log; cat $O/pytest_gpu.log; cat $O/bench_guided.json
