mkdir -p gpurun_out
lscpu | head -25 > gpurun_out/lscpu_r02c.txt; nproc >> gpurun_out/lscpu_r02c.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed_upload" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_cpp_mirror.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/upload_bench.py --gb 5 --threads 4,8,12,15,24,31 2>&1 | tail -12
for v in base add3_1 add3_2 add3_3 base; do LM_B200_LIB=tools/_variants/lib_$v.so timeout 120 python tools/kbench.py --bp 5000000000 --reps 8 --tag $v 2>&1 | tail -1 | cut -c1-150; done
