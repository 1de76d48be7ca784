#!/bin/bash
# N-GPU bench line (torchrun, one rank per GPU) incl. the copy-only ceiling; N from $NGPU (default 8)
N=${NGPU:-8}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi topo -m > gpurun_out/r2_topo_n${N}.txt 2>&1
lscpu | grep -i "numa\|model name\|socket\|^CPU(s)" > gpurun_out/r2_lscpu.txt 2>&1
for cs in ${COPY_STREAMS:-2}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --copy-streams $cs > gpurun_out/r2_bench_n${N}_cs${cs}.json 2> gpurun_out/r2_bench_n${N}_cs${cs}.err
echo "rc $?"; tail -3 gpurun_out/r2_bench_n${N}_cs${cs}.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_n${N}_cs${cs}.json"))
    print("N=${N} cs=${cs} value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ceiling", d["copy_ceiling"])
except Exception as e: print("no line", e)
PY
done
