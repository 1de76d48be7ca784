#!/bin/bash
# usage: gpu_ncu.sh <kernel-regex> <tag> [skip] [count]
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-0} -c ${4:-4} -o gpurun_out/prof_$2 -f \
   python bench.py --per-gpu 64 --chunk 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$2.log | cut -c1-300
