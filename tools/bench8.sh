#!/bin/bash
# The driver's N-GPU launch of bench.py (ours + reference arm) and the real multi-GPU check of the element-decomposed ADI.
# usage: bash tools/bench8.sh N TAG
N=${1:-8}; TAG=${2:-r2}
mkdir -p gpurun_out
export DD_VARIANTS="POP_ROW2_DC=4;POP_ROW2_HATMODE=0"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/dd_check.py > gpurun_out/${TAG}_dd${N}.log 2>&1
echo "dd_check rc=$?"; grep "^world" gpurun_out/${TAG}_dd${N}.log | cut -c1-260
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench rc=$?"; tail -c 900 gpurun_out/${TAG}_bench_n${N}.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus $N > gpurun_out/${TAG}_bench_n${N}_reference.json 2> gpurun_out/${TAG}_bench_n${N}_reference.err
echo "reference rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_n${N}_reference.json
