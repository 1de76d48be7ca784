#!/bin/bash
mkdir -p gpurun_out
for cfg in "32 4" "32 6" "32 8" "48 4" "24 6"; do
  set -- $cfg
  timeout 600 python bench.py --per-gpu 1024 --chunk $1 --streams $2 --steps 2 --warmup 2 --no-cpu-baseline --no-roofline2048 > gpurun_out/bench_t_c$1_s$2.json 2> gpurun_out/bench_t_c$1_s$2.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_t_c$1_s$2.json') if l.startswith('{')][-1])
    print('chunk=$1 streams=$2', round(d['value']), 'img/s e2e', round(d['e2e']['value']), d['check'])
except Exception as e:
    print('chunk=$1 streams=$2 failed', e); print(open('gpurun_out/bench_t_c$1_s$2.err').read()[-600:])
PY
done
