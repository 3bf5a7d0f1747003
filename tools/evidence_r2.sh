#!/bin/bash
# Round-2 evidence on one GPU box: ncu full captures of the kernels that carry the ADI, launch list of the bench.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "bitwise_reproducible" > gpurun_out/r2_repro.log 2>&1; tail -2 gpurun_out/r2_repro.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_row2 -s 2 -c 1 -f -o gpurun_out/r2_ncu_row2 python tools/row2_check.py > gpurun_out/r2_ncu_row2.log 2>&1; echo "ncu row2 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused2 -s 2 -c 1 -f -o gpurun_out/r2_ncu_fused2_mul python tools/prof_adi.py 8191 > gpurun_out/r2_ncu_fused2_mul.log 2>&1; echo "ncu fused2 mul rc=$?"
POP_ROW2=0 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_row_bwd|k_row_hats|k_row_fwd_mul|k_row_hat_mul" -s 4 -c 4 -f -o gpurun_out/r2_ncu_rows_twopass python tools/row2_check.py > gpurun_out/r2_ncu_rows_twopass.log 2>&1; echo "ncu two-pass rows rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
