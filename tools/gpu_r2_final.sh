#!/bin/bash
# Round-2 final evidence session (one B200): parity tests, both bench arms, every bench workload, phase tables, the ncu launch
# list, ncu --set full of every hot kernel at config-2 size and of block 0's kernels at config-3 size (pages exported on the
# box: the .ncu-rep files exceed the return limit), one complete T = 1000 sampling run.  usage: tools/gpu_r2_final.sh [outdir]
set -u
O=gpurun_out/${1:-r2final}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8) > $O/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py > $O/bench_guided.json 2> $O/bench_guided.err
timeout 300 python bench.py --no-graph --no-cpu-baseline > $O/bench_guided_eager.json 2> $O/bench_guided_eager.err
MDB_OVERLAP=0 timeout 300 python bench.py --no-graph --no-cpu-baseline > $O/bench_guided_eager_nooverlap.json 2> $O/bench_guided_eager_nooverlap.err
timeout 300 python bench.py --workload unguided > $O/bench_unguided.json 2> $O/bench_unguided.err
timeout 300 python bench.py --simple --no-cpu-baseline > $O/bench_simple.json 2> $O/bench_simple.err
timeout 300 python bench.py --simple --no-graph --no-cpu-baseline > $O/bench_simple_eager.json 2> $O/bench_simple_eager.err
timeout 300 python bench.py --workload train_fwd --no-cpu-baseline > $O/bench_train_fwd.json 2> $O/bench_train_fwd.err
timeout 300 python bench.py --workload unguided --batch 1024 --no-cpu-baseline > $O/bench_unguided_b1024.json 2> $O/bench_unguided_b1024.err
timeout 600 python bench.py --workload unguided --batch 8192 --max-size 29 --no-cpu-baseline > $O/bench_unguided_qm9_b8192.json 2> $O/bench_unguided_qm9_b8192.err
timeout 600 python tools/full_sample_run.py $O/full_sample_guided.json > $O/full_sample_guided.log 2>&1
timeout 300 python tools/full_sample_run.py $O/full_sample_unguided.json --unguided > $O/full_sample_unguided.log 2>&1
timeout 300 python tools/tc_phase_times.py > $O/phase_times_fwd16.txt 2>&1
timeout 300 python tools/tc_phase_times_bwd.py > $O/phase_times_bwd16.txt 2>&1
timeout 300 python tools/tc_phase_times_ffn.py > $O/phase_times_ffn2.txt 2>&1
export MDB_OVERLAP=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv \
   python bench.py --workload guided --no-graph --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
for k in tc_nodeblock_fwd16 tc_nodeblock_bwd16 tc_bondffn_fwd2 tc_bondffn_bwd3 tc_edge_d tc_node_kernel tc_bwd_node tc_edge_tail_bwd ^node_kernel transition_step; do
  n=${k#^}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 -c 1 -f -o $O/full_$n \
     python bench.py --workload guided --no-graph --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_$n.log 2>&1
  ncu -i $O/full_$n.ncu-rep --page raw --csv > $O/raw_$n.csv 2>/dev/null
  ncu -i $O/full_$n.ncu-rep --page source --csv --print-source sass > $O/src_sass_$n.csv 2>/dev/null
  ncu -i $O/full_$n.ncu-rep --page details > $O/details_$n.txt 2>/dev/null
  gzip -f $O/src_sass_$n.csv
  rm -f $O/full_$n.ncu-rep
done
# BASELINE config 3 (train_MolDiff.yml, B = 1024): block 0's kernels of the get_loss forward
for k in tc_bondffn_fwd2 tc_nodeblock_fwd16 tc_edge_d tc_node_kernel; do
  timeout 400 ncu --set full --clock-control none -k regex:$k --launch-skip 0 -c 1 -f -o $O/full_cfg3_$k \
     python bench.py --workload train_fwd --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_cfg3_$k.log 2>&1
  ncu -i $O/full_cfg3_$k.ncu-rep --page raw --csv > $O/raw_cfg3_$k.csv 2>/dev/null
  ncu -i $O/full_cfg3_$k.ncu-rep --page details > $O/details_cfg3_$k.txt 2>/dev/null
  rm -f $O/full_cfg3_$k.ncu-rep
done
unset MDB_OVERLAP
cat $O/pytest_gpu.log; du -sh $O
for f in guided guided_eager guided_eager_nooverlap unguided simple simple_eager train_fwd unguided_b1024 unguided_qm9_b8192 reference; do echo "== $f"; head -c 420 $O/bench_$f.json; echo; done
cat $O/full_sample_guided.log | tail -2; cat $O/full_sample_unguided.log | tail -1
