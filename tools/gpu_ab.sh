#!/bin/bash
# usage: gpu_ab.sh "<legacy variants>" "<ncu kernel regex>"   (A/B pass, CSV export on the box)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
VARIANTS=${1:-"none"}; KREGEX=${2:-""}; TAG=${3:-c}
T="python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider"
timeout 900 $T > gpurun_out/pytest_$TAG.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/pytest_$TAG.log
if [ $rc -ne 0 ]; then
  grep -E "^(FAILED|ERROR)" gpurun_out/pytest_$TAG.log | head -20
  for leg in $VARIANTS; do
    [ $leg = none ] && continue
    I2S_LEGACY=$leg timeout 600 $T -x > gpurun_out/pytest_${TAG}_$leg.log 2>&1; echo "legacy=$leg rc=$?"; tail -1 gpurun_out/pytest_${TAG}_$leg.log
  done
fi
B="python bench.py --per-gpu 512 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-roofline2048"
for leg in $VARIANTS; do
  I2S_LEGACY=$leg timeout 600 $B > gpurun_out/bench_${TAG}_$leg.json 2> gpurun_out/bench_${TAG}_$leg.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${TAG}_$leg.json') if l.startswith('{')][-1])
    print('legacy=$leg', round(d['value']), 'img/s', {k:round(v['ms_per_step'],2) for k,v in d['sections'].items() if v['ms_per_step']>0.5}, d['check'])
except Exception as e:
    print('legacy=$leg failed', e); print(open('gpurun_out/bench_${TAG}_$leg.err').read()[-600:])
PY
done
if [ -n "$KREGEX" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 0 -c 24 -o /tmp/prof_$TAG -f \
     python bench.py --per-gpu 64 --chunk 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
  ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
  for k in $(echo "$KREGEX" | tr '|' ' '); do
    ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/prof_${TAG}_src_$k.csv 2>/dev/null
  done
fi
du -sh gpurun_out
# optional 4th argument: extra bench runs with several streams
if [ -n "$4" ]; then
  for stn in $4; do
    timeout 600 python bench.py --per-gpu 512 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --streams $stn > gpurun_out/bench_${TAG}_st$stn.json 2> gpurun_out/bench_${TAG}_st$stn.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${TAG}_st$stn.json') if l.startswith('{')][-1])
    print('streams=$stn', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step', d['check'])
except Exception as e:
    print('streams=$stn failed', e); print(open('gpurun_out/bench_${TAG}_st$stn.err').read()[-600:])
PY
  done
fi
