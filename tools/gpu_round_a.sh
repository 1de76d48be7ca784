#!/bin/bash
# Round-1 evidence pass: default bench (both arms), ncu launch list, ncu --set full capture of the
# dominant kernels (exported to CSV on the box: gpurun_out/ is limited to 64 MiB).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi -L > gpurun_out/smi.log
timeout 900 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_a.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_a_ref.json 2> gpurun_out/bench_a_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_a.csv \
   python bench.py --per-gpu 256 --chunk 128 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_a.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on \
   -k regex:'k_vote_peaks|k_edge_buckets|k_sobel_nms|k_median|k_gauss357|k_hysteresis|k_radius|k_classify|k_circles_finish|k_mask|k_line_vote|k_grey' \
   -s 0 -c 64 -o /tmp/prof_a -f \
   python bench.py --per-gpu 64 --chunk 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_a.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_a.log | cut -c1-300
ncu -i /tmp/prof_a.ncu-rep --page raw --csv > gpurun_out/prof_a_raw.csv 2>/dev/null
for k in k_vote_peaks k_edge_buckets k_sobel_nms k_median k_gauss357 k_hysteresis k_radius; do
  ncu -i /tmp/prof_a.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/prof_a_src_$k.csv 2>/dev/null
done
du -sh gpurun_out; ls -la gpurun_out | head -40
