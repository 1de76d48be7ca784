#!/bin/bash
# ncu evidence: launch list of one timed step + one full capture of every kernel of the path
# (64 images / 512 maps per launch), exported to CSV on the box.
TAG=${TAG:-r2}
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
ARGS="--no-cpu-baseline --no-e2e --no-extras --no-grey"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --per-gpu 64 --steps 1 --warmup 1 --streams 1 $ARGS > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:"${KERNELS:-k_vote_peaks|k_edge_list|k_canny_roll|k_median|k_gauss357_roll|k_hysteresis_list|k_radius|k_classify|k_enhance|k_stack|k_line_peaks|k_circles_finish|k_mask|k_line_vote}" \
   -s 0 -c ${COUNT:-60} -o /tmp/prof_${TAG} -f \
   python bench.py --per-gpu 64 --chunk 64 --streams 1 --steps 1 --warmup 1 $ARGS > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/${TAG}_ncu.log | cut -c1-200
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw.csv 2>/dev/null
for k in ${SRC_KERNELS:-k_vote_peaks k_edge_list k_canny_roll k_median k_gauss357_roll k_radius k_circles_finish}; do
  ncu -i /tmp/prof_${TAG}.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/${TAG}_src_$k.csv 2>/dev/null
done
du -sh gpurun_out
