#!/bin/bash
# gpurun batch AA: where does the FC1 + QuickGELU epilogue's time go?  Kernel alone with parts of the epilogue removed
# (MMB_GEMM_DEBUG: 1 = no second TMA store, 2 = no activation math, 4 = no TMA stores, 8 = no fence / group barrier).
mkdir -p gpurun_out
for d in 0 1 2 4 8 12 14; do
  echo "== MMB_GEMM_DEBUG=$d"
  MMB_GEMM_DEBUG=$d timeout 300 python scripts/gemm_bench.py 2>&1 | grep -E "fc1_fwd\+act|fc2_dgrad|qkv_fwd|txt" | head -5 | cut -c1-90
done > gpurun_out/r2aa_gemm_debug.log 2>&1
cat gpurun_out/r2aa_gemm_debug.log
