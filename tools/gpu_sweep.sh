#!/bin/bash
OUT=gpurun_out/sweep.log
: > $OUT
for lib in tools/_variants/lib_*.so; do
  LM_B200_LIB=$lib timeout 120 python tools/kbench.py --bp 5000000000 --reps 5 --boot 65536 --tag $(basename $lib .so) >> $OUT 2>&1
done
LM_B200_LIB=tools/_variants/lib_pl27_ns2.so timeout 120 python tools/kbench.py --bp 5000000000 --reps 5 --boot 65536 --fmode 2 --rows 5 --tag pl27_pair5 >> $OUT 2>&1
LM_B200_LIB=tools/_variants/lib_pl27_ns2.so timeout 120 python tools/kbench.py --bp 5000000000 --reps 5 --boot 65536 --fmode 1 --rows 8 --tag pl27_single8 >> $OUT 2>&1
cut -c1-100,200-330 $OUT
