#!/bin/bash
# compute-sanitizer passes over the smallest parity cases of the kernels that rest on mailbox / mbarrier protocols
# (k_fused2 with clusters of 1, 4, 8; k_fused_solve; k_row1; k_row2; the element-decomposed row kernels).
# usage: tools/sanitize.sh TOOL [TAG]     TOOL = memcheck | racecheck | synccheck | initcheck ; output: gpurun_out/sanitizer_TOOL_TAG.log
TOOL=${1:-memcheck}
TAG=${2:-r2}
mkdir -p gpurun_out
SEL='test_solve_staged_and_fused_vs_oracle or test_fused_solve_mul_axpy or (test_single_pass_tma_row_half_step and (dirichlet-8-12 or c1-8-9-4 or dirichlet-16-9-8)) or (test_one_pass_row_half_step and dirichlet-8-12-4) or (test_element_decomposed_row_half_step_and_adi and dirichlet-6-10) or test_rdiv_one_pass_rows'
timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $TOOL --error-exitcode 66 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer_${TOOL}_$TAG.log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_${TOOL}_$TAG.log | sort | uniq -c | sort -rn | head -12
