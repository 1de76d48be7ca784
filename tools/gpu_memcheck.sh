#!/bin/bash
# compute-sanitizer memcheck over the smoke invocation and the odd-shape / saturated primitive tests
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; grep -E "ERROR SUMMARY|smoke ok" gpurun_out/memcheck_smoke.log | tail -3
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "odd_shapes or saturated or ragged or long_chain" > gpurun_out/memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_tests.log | tail -4
grep -E "Invalid|out of bounds|misaligned" gpurun_out/memcheck_smoke.log gpurun_out/memcheck_tests.log | head -10
