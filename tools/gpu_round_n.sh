#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
bash tools/gpu_round_c.sh "none radius32" "" n
timeout 900 python bench.py > gpurun_out/bench_n_full.json 2> gpurun_out/bench_n_full.err; echo "full bench rc=$?"; tail -2 gpurun_out/bench_n_full.err; cut -c1-300 gpurun_out/bench_n_full.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n_full.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],3), 'config2', d['roofline_config2'])
PY
