mkdir -p gpurun_out
timeout 300 python tools/stripe_bench.py 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stripe_kernel -s 1 -c 1 -f -o gpurun_out/stripe_r02d python tools/stripe_bench.py > gpurun_out/ncu_stripe_r02d.log 2>&1
tail -1 gpurun_out/ncu_stripe_r02d.log
