#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_smoke.py > gpurun_out/r2_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|clean:|ragged:|stage calls" gpurun_out/r2_$tool.txt | head
done
