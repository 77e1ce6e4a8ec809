#!/bin/bash
# gpurun batch N: phase trace of the fused backward workers
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2n_build.log 2>&1
timeout 200 python scripts/attn_bwd_trace.py > gpurun_out/r2n_bwd_trace.log 2>&1
cat gpurun_out/r2n_bwd_trace.log
