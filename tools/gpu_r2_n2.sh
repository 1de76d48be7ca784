#!/bin/bash
# N=2 record under torchrun (NCCL all-gather of the records), both arms; plus the wall-clock of the default N=1 run
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for impl in ours reference; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus 2 --steps 10 --warmup 3 --impl $impl > gpurun_out/r2_n2_$impl.out 2> gpurun_out/r2_n2_$impl.err
  echo "$impl rc $?"; tail -1 gpurun_out/r2_n2_$impl.out | cut -c1-240
done
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_n1_default.json 2> gpurun_out/r2_n1_default.err
echo "default N=1 rc $? wall $(( $(date +%s) - S )) s"
