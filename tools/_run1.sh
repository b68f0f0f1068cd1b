timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compat.py tests/test_cpp_mirror.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/kbench.py --bp 1000000000 --reps 5 --prune 0 --tag exact_prefix 2>&1 | tail -1 | cut -c1-110
timeout 120 python tools/kbench.py --bp 1000000000 --reps 5 --prune 0 --m 12 --tag exact_m12 2>&1 | tail -1 | cut -c1-110
timeout 120 python tools/kbench.py --bp 1000000000 --reps 5 --prune 0 --m 32 --tag exact_m32 2>&1 | tail -1 | cut -c1-110
