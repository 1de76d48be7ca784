#!/bin/bash
# last N=1 record of the round: parity tests, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_o.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_o.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_o.log
timeout 900 python bench.py > gpurun_out/bench_o_n1.json 2> gpurun_out/bench_o_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_o_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_o_n1_ref.json 2> gpurun_out/bench_o_n1_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_o_n1.json') if l.startswith('{')][-1])
r=json.loads([l for l in open('gpurun_out/bench_o_n1_ref.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],3), 'config2', round(d['roofline_config2']['frac'],3), 'ref', round(r['value'],1), d['clocks'], d['check'])
print({k:round(v['ms_per_step'],2) for k,v in d['sections'].items() if v['ms_per_step']>1})
PY
