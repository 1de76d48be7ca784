#!/bin/bash
# first GPU call: diagnosis + parity tests + smoke + a small bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.log 2>&1; nproc >> gpurun_out/smi.log; free -g >> gpurun_out/smi.log
python -c "import cv2, sklearn; print('cv2', cv2.__version__)" >> gpurun_out/smi.log 2>&1
timeout 600 python tools/gpu_diag.py ex9 synth1 > gpurun_out/diag1.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag1.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke1.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke1.log
timeout 900 python bench.py --per-gpu 128 --steps 2 --warmup 1 --cpu-images 64 > gpurun_out/bench1.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench1.log
tail -30 gpurun_out/diag1.log; tail -40 gpurun_out/pytest1.log; tail -5 gpurun_out/smoke1.log; tail -5 gpurun_out/bench1.log
