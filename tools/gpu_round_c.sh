#!/bin/bash
# GPU session C (final evidence of a round): parity tests, both bench arms, ncu launch list, ncu --set full of the hot
# kernels (pages exported on the box: the .ncu-rep files exceed the return limit), phase timings.
set -u
O=gpurun_out/${1:-c}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8) > $O/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py > $O/bench_guided.json 2> $O/bench_guided.err
timeout 300 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided.json 2> $O/bench_unguided.err
timeout 300 python tools/tc_phase_times.py > $O/phase_times_fwd16.txt 2>&1
timeout 300 python tools/tc_phase_times_bwd.py > $O/phase_times_bwd16.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv \
   python bench.py --workload guided --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
for k in tc_nodeblock_fwd16 tc_nodeblock_bwd16 tc_bondffn_fwd tc_bondffn_bwd tc_edge_d tc_node_kernel bwd_node_kernel bwd_edge_tail transition_step; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 -c 1 -f -o $O/full_$k \
     python bench.py --workload guided --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_$k.log 2>&1
  ncu -i $O/full_$k.ncu-rep --page raw --csv > $O/raw_$k.csv 2>/dev/null
  ncu -i $O/full_$k.ncu-rep --page source --csv --print-source sass > $O/src_sass_$k.csv 2>/dev/null
  ncu -i $O/full_$k.ncu-rep --page details > $O/details_$k.txt 2>/dev/null
  gzip -f $O/src_sass_$k.csv
  rm -f $O/full_$k.ncu-rep
done
cat $O/pytest_gpu.log; du -sh $O
