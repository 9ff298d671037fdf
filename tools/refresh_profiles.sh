#!/bin/bash
# Copies the outputs of a tools/gpu_r2_final.sh session (gpurun_out/<dir>) into profiles/r02_* and digests the ncu pages.
# usage: tools/refresh_profiles.sh r2final3
set -eu
O=gpurun_out/$1; P=profiles
for f in guided guided_eager guided_eager_nooverlap unguided simple simple_eager train_fwd reference; do cp $O/bench_$f.json $P/r02_bench_$f.json; done
cp $O/bench_unguided_b1024.json $P/r02_bench_unguided_b1024_config3_graph.json
cp $O/bench_unguided_qm9_b8192.json $P/r02_bench_unguided_qm9_b8192_config5_graph.json
cp $O/full_sample_guided.json $P/r02_full_sample_T1000_guided.json; cp $O/full_sample_unguided.json $P/r02_full_sample_T1000_unguided.json
cp $O/phase_times_fwd16.txt $P/r02_phase_times_fwd16.txt; cp $O/phase_times_bwd16.txt $P/r02_phase_times_bwd16.txt; cp $O/phase_times_ffn2.txt $P/r02_phase_times_ffn2.txt
cp $O/pytest_gpu.log $P/r02_pytest_gpu_session.log; cp $O/smi.txt $P/r02_smi.txt
python tools/launch_shares.py $O/launches.csv $P/r02_launch_shares.csv "MDB_OVERLAP=0 python bench.py --workload guided --no-graph --steps 1 --warmup 3 --no-cpu-baseline" > /dev/null
KS="tc_nodeblock_fwd16 tc_nodeblock_bwd16 tc_bondffn_fwd2 tc_bondffn_bwd3 tc_edge_d tc_node_kernel tc_bwd_node tc_edge_tail_bwd node_kernel transition_step"
(for k in $KS; do echo "################ $k (config 2: B=256, N=6286, E=157102; guided step; MDB_OVERLAP=0, eager launches)"; cat $O/details_$k.txt; done) > $P/r02_ncu_details.txt
(for k in tc_bondffn_fwd2 tc_nodeblock_fwd16 tc_edge_d tc_node_kernel; do echo "################ $k, block 0 of MolDiff.get_loss at BASELINE config 3 (train_MolDiff.yml, B=1024: N=24947, E=613694)"; cat $O/details_cfg3_$k.txt; done) > $P/r02_ncu_details_config3_block0.txt
python tools/ncu_metrics.py $P/r02_ncu_metrics.json $(for k in $KS; do echo $O/raw_$k.csv; done) | cut -c1-200
python tools/sass_hist.py > $P/r02_sass_opcodes.txt
