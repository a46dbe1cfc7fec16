#!/bin/bash
# one `ncu --set full` capture per kernel family (reports in gpurun_out/, summaries by tools/ncu_summary.py)
run() { ncu --set full --import-source on --clock-control none -k "regex:$2" -c 1 -s "${3:-2}" -o gpurun_out/k_$1 -f python tools/prof_kernels.py $1 2>&1 | tail -1; }
run large condense_large_kernel
run warp condense_warp
run warp78 condense_warp
run cw33 condense_cw_kernel
run gather gather_nzval
run backsub_dmma backsub_dmma_kernel
run backsub_factors backsub_factors_kernel
run batched_solve batched_solve_kernel
run expand expand_records_kernel
