#!/bin/bash
# A/B pass for the second-generation kernels (rolling Gaussian / Sobel+NMS, compacted vote).
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
T="python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider"
timeout 900 $T > gpurun_out/pytest_b.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/pytest_b.log
if [ $rc -ne 0 ]; then
  grep -E "^(FAILED|ERROR)" gpurun_out/pytest_b.log | head -20
  for leg in gauss sobel vote; do
    I2S_LEGACY=$leg timeout 600 $T -x > gpurun_out/pytest_b_$leg.log 2>&1; echo "legacy=$leg rc=$?"; tail -1 gpurun_out/pytest_b_$leg.log
  done
fi
B="python bench.py --per-gpu 512 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e"
for leg in none gauss sobel vote gauss,sobel,vote; do
  I2S_LEGACY=$leg timeout 600 $B > gpurun_out/bench_b_$leg.json 2> gpurun_out/bench_b_$leg.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_b_$leg.json') if l.startswith('{')][-1])
    print('legacy=$leg', round(d['value']), 'img/s', {k:round(v['ms_per_step'],2) for k,v in d['sections'].items() if v['ms_per_step']>0.5}, d['check'])
except Exception as e:
    print('legacy=$leg failed', e); print(open('gpurun_out/bench_b_$leg.err').read()[-600:])
PY
done
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:'k_vote_peaks2|k_canny_roll|k_gauss357_roll|k_edge_buckets' -s 0 -c 12 -o /tmp/prof_b -f \
   python bench.py --per-gpu 64 --chunk 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_b.ncu-rep --page raw --csv > gpurun_out/prof_b_raw.csv 2>/dev/null
for k in k_vote_peaks2 k_canny_roll k_gauss357_roll; do
  ncu -i /tmp/prof_b.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/prof_b_src_$k.csv 2>/dev/null
done
du -sh gpurun_out
