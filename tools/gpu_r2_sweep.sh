#!/bin/bash
# stream / chunk sweep of the device-resident headline rate
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for cfg in ${CFGS:-"8 32" "16 32" "8 16" "16 16" "4 64" "12 24" "6 48"}; do
  set -- $cfg
  timeout 300 python bench.py --steps 3 --warmup 2 --streams $1 --chunk $2 --no-cpu-baseline --no-e2e --no-extras --no-grey 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('streams $1 chunk $2 value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'single-stream', round(d['sections_pass']['ms_per_step'],2))"
done | tee gpurun_out/r2_sweep.txt
