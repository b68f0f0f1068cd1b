#!/bin/bash
# run tools/kbench.py for every experiment library under gpurun_out/variants/
OUT=gpurun_out/variants.log
: > $OUT
for lib in tools/_variants/lib_*.so; do
  tag=$(basename $lib .so)
  LM_B200_LIB=$lib timeout 120 python tools/kbench.py --bp 5000000000 --reps 6 --tag $tag "$@" >> $OUT 2>&1
done
cat $OUT
