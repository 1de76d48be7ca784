#!/bin/bash
# Evidence pass of round 2 on one B200: parity suite, smoke, the default bench line (both arms), sanitizers,
# the ncu launch list of one timed step and one full capture of every kernel (exported to CSV on the box).
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
nvidia-smi -L > gpurun_out/r2_smi.txt
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_n1.json 2> gpurun_out/r2_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_n1.err; tail -1 gpurun_out/r2_n1.json | cut -c1-260
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_n1_ref.json 2> gpurun_out/r2_n1_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/r2_n1_ref.json | cut -c1-200
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_smoke.py > gpurun_out/r2_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_$tool.txt | head -3
done
TAG=r2 COUNT=70 bash tools/gpu_r2_ncu.sh
