#!/bin/bash
# ncu --set full of ONE kernel (regex $1, launch-skip $2), pages exported on the box (reports exceed the return limit)
set -u
k=$1; skip=${2:-1}; O=gpurun_out/ncu; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 -f -o $O/full_$k \
   python bench.py --workload guided --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_$k.log 2>&1
ncu -i $O/full_$k.ncu-rep --page raw --csv > $O/raw_$k.csv 2>/dev/null
ncu -i $O/full_$k.ncu-rep --page source --csv --print-source sass > $O/src_sass_$k.csv 2>/dev/null
ncu -i $O/full_$k.ncu-rep --page details > $O/details_$k.txt 2>/dev/null
gzip -f $O/src_sass_$k.csv
rm -f $O/full_$k.ncu-rep
tail -3 $O/ncu_$k.log
