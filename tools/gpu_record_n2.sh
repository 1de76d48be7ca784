#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_p_n2.json 2> gpurun_out/bench_p_n2.err; echo "n2 rc=$?"
tail -2 gpurun_out/bench_p_n2.err; grep '^{' gpurun_out/bench_p_n2.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_p_n2_ref.json 2> gpurun_out/bench_p_n2_ref.err; echo "n2 ref rc=$?"
grep '^{' gpurun_out/bench_p_n2_ref.json | cut -c1-200
