#!/bin/bash
# the default bench line (all blocks) + peak device memory of the run
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits; sleep 1; done ) > gpurun_out/r2_mem.log 2>/dev/null &
MON=$!
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc $?"
kill $MON
tail -3 gpurun_out/r2_bench_b.err; echo "peak MiB: $(sort -n gpurun_out/r2_mem.log | tail -1)"
