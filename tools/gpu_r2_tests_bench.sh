#!/bin/bash
# First GPU pass of a change: the parity suite, then a short N=1 bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/r2_pytest.log
tail -40 gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps ${STEPS:-3} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
echo "bench rc $?"; tail -5 gpurun_out/r2_bench_a.err; head -c 3000 gpurun_out/r2_bench_a.json
