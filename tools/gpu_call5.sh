#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest5.log
tail -15 gpurun_out/pytest5.log
timeout 900 python bench.py --per-gpu 256 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench5.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench5.log
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench5.log') if l.startswith('{')][-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'check',d['check'])
    for k,v in d['sections'].items(): print(f"  {k:16s} {v['ms_per_step']:8.3f} ms")
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench5.log').read()[-2000:])
PY
bash tools/gpu_ncu.sh "k_hysteresis" r1_e_hyst 8 1
bash tools/gpu_ncu.sh "k_median|k_sobel_nms" r1_e_med 0 5
