#!/bin/bash
# N=1 default bench (both arms) + N=2 torchrun bench (both arms) for the record
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_g.log
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_g.log
timeout 900 python bench.py > gpurun_out/bench_g_n1.json 2> gpurun_out/bench_g_n1.err; echo "n1 rc=$?"; tail -2 gpurun_out/bench_g_n1.err; cut -c1-400 gpurun_out/bench_g_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_g_n1_ref.json 2> gpurun_out/bench_g_n1_ref.err; echo "n1 ref rc=$?"; cut -c1-200 gpurun_out/bench_g_n1_ref.json
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_g_n2.json 2> gpurun_out/bench_g_n2.err; echo "n2 rc=$?"
  tail -3 gpurun_out/bench_g_n2.err; cut -c1-400 gpurun_out/bench_g_n2.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_g_n2_ref.json 2> gpurun_out/bench_g_n2_ref.err; echo "n2 ref rc=$?"
  cut -c1-200 gpurun_out/bench_g_n2_ref.json
fi
