#!/bin/bash
# Real 8-GPU run of the element-decomposed ADI (tools/dd_check.py): correctness at small sizes, J = 90 timing, the same under
# other switches of the kernels, and the phase trace of k_row2 on ranks 0 and 4.
mkdir -p gpurun_out
export DD_TRACE=1
export DD_VARIANTS="POP_ROW2_HATMODE=0;POP_F_EVICT_FIRST=0;POP_ROW2_DC=8"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/dd_check.py > gpurun_out/dd8_variants.log 2>&1
echo "rc=$?"; grep -v "^W\|warn" gpurun_out/dd8_variants.log | cut -c1-420 | tail -60
