for v in ilv12 b4096 b2048 b1024; do echo "$v"; LM_B200_LIB=tools/_variants/lib_$v.so timeout 300 python tools/config2_bench.py 2>&1 | tail -1 | cut -c100-230; done
