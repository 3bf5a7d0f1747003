#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu full capture of the headline kernel.
# usage: tools/round_check.sh TAG   (outputs under gpurun_out/)
TAG=${1:-r1z}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused2 -s 3 -c 1 -f -o gpurun_out/prof_bench_fused2_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-adi > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k_analysis_2d|k_weakform" -c 12 -f -o gpurun_out/prof_rhs_$TAG \
    env RHS_CPU=0 python tools/prof_rhs.py > gpurun_out/ncu_rhs_$TAG.log 2>&1; echo "ncu rhs rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_row1 -s 6 -c 1 -f -o gpurun_out/prof_row1_$TAG \
    python tools/row1_check.py 128 64 > gpurun_out/ncu_row1_$TAG.log 2>&1; echo "ncu row1 rc=$?"
timeout 100 python tools/prof_rhs.py > gpurun_out/prof_rhs_$TAG.log 2>&1
timeout 100 python tools/row1_check.py 128 64 --adi > gpurun_out/row1_check_$TAG.log 2>&1
head -c 1500 gpurun_out/bench_$TAG.json
