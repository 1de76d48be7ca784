#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
bash tools/gpu_round_c.sh "none mask vote" "k_vote_peaks2|k_hysteresis|k_mask" h
for cfg in "64 2" "64 3" "256 2" "128 3"; do
  set -- $cfg
  timeout 600 python bench.py --per-gpu 1024 --chunk $1 --streams $2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_h_c$1_s$2.json 2> gpurun_out/bench_h_c$1_s$2.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_h_c$1_s$2.json') if l.startswith('{')][-1])
    print('chunk=$1 streams=$2', round(d['value']), 'img/s e2e', round(d['e2e']['value']), d['check'])
except Exception as e:
    print('chunk=$1 streams=$2 failed', e); print(open('gpurun_out/bench_h_c$1_s$2.err').read()[-600:])
PY
done
