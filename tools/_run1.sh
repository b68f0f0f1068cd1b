mkdir -p gpurun_out
LM_B200_LIB=tools/_variants/lib_async.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -s 2 -c 1 -f -o gpurun_out/prune_async2 python tools/kbench.py --bp 5000000000 --reps 4 --tag async > gpurun_out/ncu_async.log 2>&1
tail -1 gpurun_out/ncu_async.log | cut -c1-150
