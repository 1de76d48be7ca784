#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest6.log
for c in 32 64 128 256; do
  timeout 600 python bench.py --per-gpu 512 --chunk $c --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/sweep_$c.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/sweep_$c.log') if l.startswith('{')][-1])
print('chunk $c value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],1))
PY
done
timeout 1200 python bench.py > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/bench_r1_b_ref.json 2> gpurun_out/bench_r1_b_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_r1_b.json | cut -c1-1800; echo; cat gpurun_out/bench_r1_b_ref.json | cut -c1-900
