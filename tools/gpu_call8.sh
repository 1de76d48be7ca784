#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest8.log
for st in 1 2 3; do
  timeout 600 python bench.py --per-gpu 512 --chunk 128 --streams $st --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/streams_$st.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/streams_$st.log') if l.startswith('{')][-1])
print('streams $st value',round(d['value']),'ms/step',round(d['ms_per_step'],1), {k:round(v['ms_per_step'],2) for k,v in d['sections'].items() if v['ms_per_step']>1})
PY
done
