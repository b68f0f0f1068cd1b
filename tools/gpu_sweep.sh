#!/bin/bash
# run tools/kbench.py for every experiment library under tools/_variants/
OUT=gpurun_out/sweep.log
: > $OUT
for lib in tools/_variants/lib_*.so; do
  LM_B200_LIB=$lib timeout 45 python tools/kbench.py --bp 5000000000 --reps 6 --tag $(basename $lib .so) "$@" >> $OUT 2>&1
done
cut -c1-100,200-330 $OUT
