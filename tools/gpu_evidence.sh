#!/bin/bash
# Final evidence pass of the round: parity tests, smoke, default bench (both arms), launch list,
# full ncu capture exported to CSV on the box (gpurun_out/ is limited to 64 MiB), 2048^2 workload.
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/prof_*_src_*.csv
nvidia-smi -L > gpurun_out/smi_final.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final_n1.err; cut -c1-300 gpurun_out/bench_final_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_n1_ref.json 2> gpurun_out/bench_final_n1_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_final_n1_ref.json
timeout 600 python bench.py --workload synth2048 --per-gpu 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_final_2048.json 2> gpurun_out/bench_final_2048.err; echo "2048 rc=$?"; tail -2 gpurun_out/bench_final_2048.err; cut -c1-300 gpurun_out/bench_final_2048.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv \
   python bench.py --per-gpu 128 --steps 1 --warmup 1 --streams 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_final.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on \
   -k regex:'k_vote_peaks2|k_edge_buckets16|k_canny_roll|k_median|k_gauss357_roll|k_hysteresis|k_radius|k_classify|k_circles_finish|k_mask|k_line_vote|k_grey' \
   -s 0 -c 70 -o /tmp/prof_final -f \
   python bench.py --per-gpu 64 --chunk 64 --streams 1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_final.log | cut -c1-200
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
for k in k_vote_peaks2 k_edge_buckets16 k_canny_roll k_median k_gauss357_roll k_radius; do
  ncu -i /tmp/prof_final.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/prof_final_src_$k.csv 2>/dev/null
done
du -sh gpurun_out
