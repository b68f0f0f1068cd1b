bash tools/gpu_round.sh r02_final
timeout 300 python tools/upload_bench.py --gb 5 --threads 15 --reps 5 > gpurun_out/upload_r02_final.txt 2>&1; tail -5 gpurun_out/upload_r02_final.txt
