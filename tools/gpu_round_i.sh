#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
bash tools/gpu_round_c.sh "none rgb5 hyst8" "k_vote_peaks2|k_radius|k_gauss357_roll" i
for cfg in "128 2" "64 3" "64 4" "32 4" "96 3"; do
  set -- $cfg
  timeout 600 python bench.py --per-gpu 1024 --chunk $1 --streams $2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_i_c$1_s$2.json 2> gpurun_out/bench_i_c$1_s$2.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_i_c$1_s$2.json') if l.startswith('{')][-1])
    print('chunk=$1 streams=$2', round(d['value']), 'img/s e2e', round(d['e2e']['value']), d['check'])
except Exception as e:
    print('chunk=$1 streams=$2 failed', e); print(open('gpurun_out/bench_i_c$1_s$2.err').read()[-600:])
PY
done
