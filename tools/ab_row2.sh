#!/bin/bash
# Same-box A/B of two builds of the library (POP_ARROWHEAD_LIB=tools/_ab/libpop_old.so is a build of an earlier commit) at the one-GPU
# and the 8-GPU (stand-in) shapes of the ADI, with the parity tests of the row kernels first.
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "remote_mailboxes or single_pass_tma or element_decomposed or adi_small or reproducible" 2>&1 | tail -3
POP_ROW2_VERBOSE=1 timeout 120 python tools/prof_dd.py 1 0 2 2>&1 | grep "pop_row2:" | sort | uniq -c | head -3
POP_DD_FAKE_PEERS=1 POP_ROW2_VERBOSE=1 timeout 120 python tools/prof_dd.py 8 1 2 2>&1 | grep "pop_row2:" | sort | uniq -c | head -3
for rep in 1 2; do
echo "== old lib"; POP_ARROWHEAD_LIB=$PWD/tools/_ab/libpop_old.so timeout 120 python tools/prof_dd.py 1 0 90 2>&1 | tail -1
echo "== new lib"; timeout 120 python tools/prof_dd.py 1 0 90 2>&1 | tail -1
echo "== new lib hat mode 1"; POP_ROW2_HATMODE=1 timeout 120 python tools/prof_dd.py 1 0 90 2>&1 | tail -1
done
for rep in 1 2; do
echo "== old w8"; POP_DD_FAKE_PEERS=1 POP_ARROWHEAD_LIB=$PWD/tools/_ab/libpop_old.so timeout 120 python tools/prof_dd.py 8 1 90 2>&1 | tail -1
echo "== new w8"; POP_DD_FAKE_PEERS=1 timeout 120 python tools/prof_dd.py 8 1 90 2>&1 | tail -1
echo "== new w8 hat mode 0"; POP_ROW2_HATMODE=0 POP_DD_FAKE_PEERS=1 timeout 120 python tools/prof_dd.py 8 1 90 2>&1 | tail -1
done
} > gpurun_out/ab_row2.log 2>&1
cat gpurun_out/ab_row2.log
