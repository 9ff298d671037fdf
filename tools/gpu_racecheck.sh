#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the small tensor-core tests.  Output: gpurun_out/san/
set -u
O=gpurun_out/san; mkdir -p $O
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 --log-file $O/racecheck.log \
  python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x --tb=short \
  -k "tc or (crossed and B16 and uncertainty) or (goldens and B4_mixed_t)" > $O/racecheck_pytest.log 2>&1
tail -3 $O/racecheck_pytest.log; grep -c "Race reported\|hazard" $O/racecheck.log; tail -6 $O/racecheck.log; grep -A6 "hazard\|Race" $O/racecheck.log | head -60
