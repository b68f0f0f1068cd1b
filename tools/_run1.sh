for b in 8192 4096 2048 1024 512; do echo "boot $b"; LM_BATCH_BOOT=$b timeout 300 python tools/config2_bench.py 2>&1 | tail -1 | cut -c100-330; done
