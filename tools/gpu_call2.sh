#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err
echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_a.csv \
   python bench.py --per-gpu 64 --chunk 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_vote|k_median|k_radius|k_sobel_nms' -s 8 -c 8 -o gpurun_out/prof_r1_a \
   python bench.py --per-gpu 64 --chunk 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_a.json | head -c 3000; tail -3 gpurun_out/ncu_list.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
