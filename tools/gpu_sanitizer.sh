#!/bin/bash
# compute-sanitizer memcheck over the tensor-core forward / backward kernels, the fused transition, decode and graph-builder
# tests (small cases; config-2-sized tests are skipped: memcheck slows the kernels ~50x).  Output: gpurun_out/san/
set -u
O=gpurun_out/san; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --log-file $O/memcheck.log \
  python -m pytest tests/test_gpu_tc.py tests/test_gpu_transition.py tests/test_gpu_decode.py tests/test_gpu_graph_build.py tests/test_gpu_parity.py \
  -m gpu -q -x --tb=short -k "not full_size and not config2 and not train_config and not training_backward and not three_training and not free_running and not sample50" \
  > $O/memcheck_pytest.log 2>&1
tail -5 $O/memcheck_pytest.log; tail -4 $O/memcheck.log
