#!/bin/bash
# sweep forced filter configurations of the product library (tools/kbench.py)
OUT=gpurun_out/sweep.log
: > $OUT
for cfg in "1 6" "1 8" "1 10" "2 4" "2 5" "2 6" "2 8" "0 0"; do
  set -- $cfg
  timeout 120 python tools/kbench.py --bp 5000000000 --reps 5 --fmode $1 --rows $2 --tag "fmode$1_r$2" >> $OUT 2>&1
done
cut -c1-100,200-330 $OUT
