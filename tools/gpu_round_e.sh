#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
bash tools/gpu_round_c.sh "none med5bits" "k_median" e
timeout 900 python bench.py > gpurun_out/bench_e_full.json 2> gpurun_out/bench_e_full.err; echo "full bench rc=$?"; tail -2 gpurun_out/bench_e_full.err; cut -c1-1200 gpurun_out/bench_e_full.json
