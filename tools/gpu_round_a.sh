#!/bin/bash
# GPU session A: parity tests, bench lines, ncu launch list, ncu --set full of the tensor-core kernels.
set -u
O=gpurun_out/a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/pytest_gpu.log
timeout 400 python bench.py --workload guided > $O/bench_guided.json 2> $O/bench_guided.err
timeout 300 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided.json 2> $O/bench_unguided.err
MDB_TC_NB16=0 timeout 300 python bench.py --workload unguided --no-cpu-baseline > $O/bench_unguided_nb8.json 2>&1
timeout 300 python tools/tc_phase_times.py > $O/phase_times.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv \
   python bench.py --workload guided --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
for k in tc_nodeblock_fwd tc_nodeblock_bwd tc_bondffn_fwd tc_bondffn_bwd tc_edge_d tc_node_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 -c 1 -f -o $O/full_$k \
     python bench.py --workload guided --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_$k.log 2>&1
  # the .ncu-rep files (15-20 MB each with sources) exceed the 64 MiB return limit: export the pages here instead
  ncu -i $O/full_$k.ncu-rep --page raw --csv > $O/raw_$k.csv 2>/dev/null
  ncu -i $O/full_$k.ncu-rep --page source --csv --print-source sass > $O/src_sass_$k.csv 2>/dev/null
  ncu -i $O/full_$k.ncu-rep --page source --csv --print-source cuda > $O/src_cuda_$k.csv 2>/dev/null
  ncu -i $O/full_$k.ncu-rep --page details > $O/details_$k.txt 2>/dev/null
  gzip -f $O/src_sass_$k.csv $O/src_cuda_$k.csv
  rm -f $O/full_$k.ncu-rep
done
du -sh $O
ls -la $O
