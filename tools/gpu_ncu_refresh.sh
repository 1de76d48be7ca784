#!/bin/bash
# ncu evidence only (launch list + one full capture), exported to CSV on the box
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/prof_*_src_*.csv gpurun_out/prof_*_raw.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_final.csv \
   python bench.py --per-gpu 64 --steps 1 --warmup 1 --streams 1 --no-cpu-baseline --no-e2e --no-roofline2048 > gpurun_out/launches_final.log 2>&1; echo "launch list rc=$?"
timeout 420 ncu --set full --clock-control none --import-source on \
   -k regex:'k_vote_peaks2|k_edge_buckets16|k_canny_roll|k_median|k_gauss357_roll|k_hysteresis_list|k_radius|k_classify|k_circles_finish|k_mask|k_line_vote|k_grey' \
   -s 0 -c 44 -o /tmp/prof_final -f \
   python bench.py --per-gpu 64 --chunk 64 --streams 1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline2048 > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/ncu_final.log | cut -c1-200
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
for k in k_vote_peaks2 k_edge_buckets16 k_canny_roll k_median k_gauss357_roll k_radius; do
  ncu -i /tmp/prof_final.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/prof_final_src_$k.csv 2>/dev/null
done
du -sh gpurun_out
