#!/bin/bash
# One gpurun call: GPU parity tests, smoke, headline bench, reference arm, ncu launch list, ncu full capture.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
lscpu | head -20 > $OUT/lscpu_$TAG.txt; nproc >> $OUT/lscpu_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -c 1500 $OUT/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-exact --no-extras --min-seconds 0 > $OUT/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -s 2 -c 1 -f -o $OUT/prune_full_$TAG \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-exact --no-extras --min-seconds 0 > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT | tail -12
